"""Time-sharding of a long IQ recording across GPUs.

FFT frames do not overlap (Scheduler.kt:254-276), so a recording of F frames is cut into
contiguous time segments, one per rank, with no halo and no data-path collective.  Only the
two N-float summaries are exchanged at the end of a pass:

  * peak hold  -- element-wise max over all rows (FftProcessor.kt:244)  -> all_reduce(MAX)
  * time average of the newest L+1 rows (AnalyzerSurface.kt:710-714) -> the newest rows live
    on the last rank(s); normally the last rank alone holds all L+1 and broadcasts its result,
    otherwise the ranks' newest rows are gathered and summed in the reference's order.

`torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is the plumbing.
"""
import torch
import torch.distributed as dist


def shard_frames(total_frames, world_size, rank):
    """Contiguous segment of rank `rank`: (first_frame, nframes); earlier ranks get the remainder."""
    base, rem = divmod(int(total_frames), int(world_size))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def tail_owner_plan(total_frames, world_size, avg_len):
    """Which ranks hold the newest avg_len+1 rows: list of (rank, rows_taken), newest first."""
    need = min(avg_len + 1, int(total_frames))
    plan = []
    for r in range(world_size - 1, -1, -1):
        if need == 0:
            break
        _, n = shard_frames(total_frames, world_size, r)
        take = min(n, need)
        if take:
            plan.append((r, take))
            need -= take
    return plan


def assemble_tail(gathered, total_frames, world_size, avg_len, fill=-9999.0):
    """gathered[r]: [avg_len+1, N] with rank r's newest rows, newest first.  Returns the global
    newest avg_len+1 rows, newest first; rows that do not exist hold -9999f (FftProcessor.kt:181)."""
    n = gathered[0].shape[-1]
    out = torch.full((avg_len + 1, n), fill, dtype=gathered[0].dtype, device=gathered[0].device)
    k = 0
    for r, take in tail_owner_plan(total_frames, world_size, avg_len):
        out[k:k + take] = gathered[r][:take]
        k += take
    return out


class ShardedSpectrum:
    """One rank's share of a time-sharded spectrum pass plus the summary reduction."""

    def __init__(self, plan, rank=0, world_size=1, group=None):
        self.plan, self.rank, self.world, self.group = plan, rank, world_size, group
        self.n, self.L = plan.fft_size, plan.avg_len
        self._packed = None

    def local_range(self, total_frames):
        return shard_frames(total_frames, self.world, self.rank)

    def process(self, iq_local, total_frames, rows_local, peaks, avg, peaks_accumulate=False, reduce=True):
        """iq_local / rows_local: this rank's segment; peaks / avg: [N] device tensors.  With
        reduce=True they hold the GLOBAL result on every rank afterwards; with reduce=False they
        hold this rank's running summaries and reduce() is called once when the recording ends
        (the peak hold and the averaged spectrum are properties of the whole recording)."""
        _, nloc = self.local_range(total_frames)
        self.plan.process(iq_local, nloc, rows=rows_local, peaks=peaks, avg=avg, peaks_accumulate=peaks_accumulate)
        if nloc == 0 and not peaks_accumulate:
            peaks.fill_(-999999.0)
        if reduce:
            self.reduce(total_frames, rows_local, peaks, avg)

    def reduce(self, total_frames, rows_local, peaks, avg):
        """Exchange the two N-float summaries in ONE collective: peaks and the averaged spectrum travel as one
        2N-float vector through all_reduce(MAX) -- the rank that holds the newest L+1 rows contributes its
        average, every other rank -inf, so the maximum IS the owner's average (bit for bit) and the peak hold
        is the element-wise maximum over all ranks (FftProcessor.kt:244).  Only when the newest L+1 rows
        straddle ranks (fewer than L+1 frames on the last rank) are the ranks' newest rows gathered and summed
        in the reference's order (AnalyzerSurface.kt:710-714)."""
        if self.world == 1:
            return
        _, nloc = self.local_range(total_frames)
        owners = tail_owner_plan(total_frames, self.world, self.L)
        n = self.n
        if len(owners) == 1 and total_frames >= self.L + 1:
            if self._packed is None or self._packed.device != peaks.device:
                self._packed = torch.empty(2 * n, dtype=torch.float32, device=peaks.device)
            packed = self._packed
            packed[:n].copy_(peaks)
            if self.rank == owners[0][0]:
                packed[n:].copy_(avg)
            else:
                packed[n:].fill_(float("-inf"))
            dist.all_reduce(packed, op=dist.ReduceOp.MAX, group=self.group)
            peaks.copy_(packed[:n])
            avg.copy_(packed[n:])
            return
        dist.all_reduce(peaks, op=dist.ReduceOp.MAX, group=self.group)
        # rare: the newest L+1 rows straddle ranks -- gather each rank's newest rows
        mine = torch.full((self.L + 1, n), -9999.0, dtype=torch.float32, device=avg.device)
        take = min(self.L + 1, nloc)
        if take:
            mine[:take] = torch.flip(rows_local[nloc - take:nloc, :n], dims=[0])
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=self.group)
        tail = assemble_tail(gathered, total_frames, self.world, self.L).contiguous()
        self.plan.ctx.average_rows(tail, 0, 1, 0, n, min(self.L + 1, total_frames), self.L, n, avg)


# ---------------------------------------------------------------------------------------------
# IQ -> audio chain (demodulation path)
# ---------------------------------------------------------------------------------------------
def shard_packets(total_samples, packet_samples, world_size, rank):
    """Contiguous run of whole packets for rank `rank`: (first_sample, nsamples).  Packetisation is part of
    the reference's semantics (per-packet mean and AGC, Demodulator.kt:293-302), so segments start on
    packet boundaries; only the recording's last packet may be short."""
    npk = -(-int(total_samples) // int(packet_samples))
    first_pk, n_pk = shard_frames(npk, world_size, rank)
    first = first_pk * packet_samples
    end = min((first_pk + n_pk) * packet_samples, int(total_samples))
    return first, max(0, end - first)


def default_halo_packets(mode):
    """Warm-up packets re-processed before a segment.  The delay lines (resampler, user / band-pass /
    audio filters) and the FM carry are exact after ONE packet; the AGC maximum of AM / SSB / CW decays by
    0.95 per packet (Demodulator.kt:290), so its memory of what came before the halo is 0.95^256 = 2e-6."""
    return 1 if mode in (2, 3) else 256  # MODE_NFM, MODE_WFM


def delay_line_span(plan):
    """Input samples that the chain's delay lines reach back: the resampler's taps per phase plus the user
    filter (27 taps), the SSB / CW band-pass (181 taps, <= 2x decimation) and the two audio decimators (9 taps at
    the demodulated rate, 13 taps after the first /2) referred to the input rate through D/I.  An upper bound --
    the halo must cover it or a rank's first audio samples differ from the sequential run."""
    banded = plan.desc.mode in (4, 5, 6)  # LSB, USB, CW
    span_q = 26 + (180 if banded else 0) + (8 + 12 * 2) * (2 if banded else 1)
    ratio = plan.decimation / plan.interpolation
    return plan.taps_per_phase + int(span_q * ratio + ratio) + 1


def halo_packets_for(plan):
    """Warm-up packets for `plan`: the mode's default, raised until it covers the delay-line span (large
    decimation with small packets, e.g. 20 Msps -> 48 kHz with 1024-sample packets)."""
    need = -(-delay_line_span(plan) // plan.desc.packet_samples)
    return max(default_halo_packets(plan.desc.mode), need)


class ShardedChain:
    """One rank's share of a time-sharded demodulation run: seek to a packet boundary one halo before the
    segment, re-process the halo (audio discarded), then the segment.  No collective on the data path: the
    audio of rank r simply follows the audio of rank r-1; `gather_layout` tells every rank where each piece
    goes."""

    def __init__(self, plan, rank=0, world_size=1, halo_packets=None, group=None):
        self.plan, self.rank, self.world, self.group = plan, rank, world_size, group
        self.packet = plan.desc.packet_samples
        self.halo_packets = halo_packets_for(plan) if halo_packets is None else int(halo_packets)
        if self.halo_packets * self.packet < delay_line_span(plan):
            raise ValueError("halo of %d packets (%d samples) is shorter than the chain's delay lines (%d samples)"
                             % (self.halo_packets, self.halo_packets * self.packet, delay_line_span(plan)))

    def segment(self, total_samples, rank=None):
        """(halo_start, first, nsamples): the rank needs input samples [halo_start, first + nsamples)."""
        first, n = shard_packets(total_samples, self.packet, self.world, self.rank if rank is None else rank)
        halo_start = max(0, first - self.halo_packets * self.packet)
        return halo_start, first, n

    def process(self, iq_from_halo, total_samples, audio):
        """iq_from_halo: raw IQ of this rank starting at its halo_start.  Writes the segment's audio to
        `audio` and returns (audio_index, n_audio): where it sits in the whole recording's audio."""
        halo_start, first, n = self.segment(total_samples)
        bps = 4 if self.plan.desc.format == 2 else 2
        index = self.plan.seek(halo_start)
        nh = first - halo_start
        if nh:
            cap = audio.numel() if hasattr(audio, "numel") else len(audio)
            if cap < self.plan.max_audio(nh):
                raise ValueError("audio buffer smaller than the halo's audio: give it max_audio(max(halo, segment))")
            index += self.plan.process(iq_from_halo[: nh * bps], nh, audio)  # overwritten by the segment's audio
        got = self.plan.process(iq_from_halo[nh * bps:(nh + n) * bps], n, audio) if n else 0
        return index, got

    def gather_layout(self, index, count):
        """all_gather of every rank's (audio_index, n_audio)."""
        if self.world == 1:
            return [(index, count)]
        mine = torch.tensor([index, count], dtype=torch.int64)
        if dist.get_backend(self.group) == "nccl":
            mine = mine.cuda()
        out = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(out, mine, group=self.group)
        return [(int(t[0]), int(t[1])) for t in out]
