// scheduler.cu -- Scheduler.run (A/analyzer/Scheduler.kt:140-298) as a batched host runtime.
//
// The reference's scheduler thread takes one packet per loop iteration from the source and fans it out:
//   :161-165  squelch debounce: squelchSatisfied -> counter = 0, else counter counts up to SQUELCH_DEBOUNCE_COUNT (50)
//   :199      recording gate  : squelchSatisfied || !onlyWhenSquelchIsSatisfied || counter < 50
//   :237-244  demodulator gate: isDemodulationActivated && (squelchSatisfied || counter < 50)
//             -> mixPacketIntoSamplePacket(packet, buffer, channelFrequency) -> demodulator queue
//   :254-276  FFT: fillPacketIntoSamplePacket APPENDS the packet's leading samples to the FFT buffer until it holds
//             fftSize samples (what does not fit is dropped), a full buffer goes to the FftProcessor
// and FftProcessor.run (:143-157) turns every frame into averageSignalStrength over the channel
// [channelFrequency - channelWidth, channelFrequency + channelWidth] (AnalyzerService.kt:344-353), from which
// squelchSatisfied = strength > squelch (AppStateRepository.kt:318-323).
//
// Here a CALL carries many packets.  The loss-free, synchronous reading of that loop is kept: buffer pools never run
// dry (no packet is dropped for lack of a buffer -- the reference drops most FFT packets only because its FFT thread is
// slow), and a frame completed by packet k has updated squelchSatisfied before packet k+1 is looked at.  Per call:
//   1. the packets' leading samples are gathered into whole FFT frames (a partial frame is carried to the next call)
//   2. ONE fused spectrum launch writes their rows into the reference's backwards ring (+ peak hold + average)
//   3. ONE launch reduces every new row to its channel strength; the strengths come back to the host
//   4. the host replays the gate sequence (squelch, debounce counter) packet by packet -- a few ns per packet
//   5. every run of consecutive packets that passed the demodulator gate goes through the IQ -> audio chain, whose
//      streaming state sees exactly the delivered packets back to back, like the reference's Demodulator thread
// With the squelch disabled the gates do not depend on the spectrum, and the chain runs on a second stream beside
// the spectrum launch.
#include <cstring>
#include <vector>

#include "capi_core.h"

using namespace rfa;

namespace {

constexpr int kSquelchDebounceCount = 50;  // Scheduler.kt:52

struct GatherSeg {
    long long src;  // byte offset into the packets
    long long dst;  // byte offset into the frame buffer
    int bytes;
    int pad;
};

// copies seg[blockIdx.y] (a few KB each) with 16-byte accesses where the alignment allows
__global__ void __launch_bounds__(256) gather_kernel(const unsigned char *__restrict__ src, unsigned char *__restrict__ dst,
                                                     const GatherSeg *__restrict__ seg, int nseg) {
    for (int s = blockIdx.y; s < nseg; s += gridDim.y) {
        const GatherSeg g = seg[s];
        const unsigned char *a = src + g.src;
        unsigned char *b = dst + g.dst;
        if ((((size_t)a | (size_t)b | (size_t)g.bytes) & 15) == 0) {
            const int n16 = g.bytes >> 4;
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x)
                reinterpret_cast<uint4 *>(b)[i] = reinterpret_cast<const uint4 *>(a)[i];
        } else {
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.bytes; i += gridDim.x * blockDim.x) b[i] = a[i];
        }
    }
}

}  // namespace

struct rfa_scheduler {
    rfa_ctx *ctx = nullptr;    // spectrum side (the caller's context)
    rfa_ctx *ctx2 = nullptr;   // chain side: a stream of its own
    rfa_scheduler_desc d{};
    rfa_spectrum_plan *plan = nullptr;
    rfa_chain *chain = nullptr;
    int bps = 2;
    int channel_width = 0;
    // FftProcessor state
    Buf ring, peaks, avg, strength;
    long long write_index = 0, history = 0;
    bool have_peaks = false;
    // Scheduler state
    bool squelch_satisfied = false;
    int debounce = 0;
    long long fill = 0;         // samples already in the partial FFT frame
    Buf partial, frames, segs, stage;
    cudaEvent_t ev = nullptr;
    long long packets_total = 0, frames_total = 0;
};

extern "C" {

int rfa_scheduler_destroy(rfa_scheduler *s) {
    if (!s) return RFA_OK;
    if (s->ctx) {
        cudaSetDevice(s->ctx->device);
        cudaStreamSynchronize(s->ctx->stream);
    }
    if (s->chain) rfa_chain_destroy(s->chain);
    if (s->plan) rfa_spectrum_plan_destroy(s->plan);
    if (s->ev) cudaEventDestroy(s->ev);
    for (Buf *b : {&s->ring, &s->peaks, &s->avg, &s->strength, &s->partial, &s->frames, &s->segs, &s->stage}) b->release();
    if (s->ctx2) rfa_ctx_destroy(s->ctx2);
    delete s;
    return RFA_OK;
}

int rfa_scheduler_create(rfa_ctx *c, const rfa_scheduler_desc *d, rfa_scheduler **out) {
    RFA_REQUIRE(c && d && out, "rfa_scheduler_create: NULL argument");
    *out = nullptr;
    RFA_REQUIRE(d->format >= RFA_FMT_S8 && d->format <= RFA_FMT_S16LE, "unknown sample format %d", d->format);
    RFA_REQUIRE(d->packet_samples > 0 && d->sample_rate > 0, "packet size and sample rate must be positive");
    RFA_REQUIRE(d->ring_rows >= 1, "the waterfall ring needs at least one row (FftProcessor.kt:104 uses 300 / 400 / 500)");
    RFA_REQUIRE(d->demodulation_mode >= RFA_MODE_OFF && d->demodulation_mode <= RFA_MODE_CW, "unknown demodulation mode %d",
                d->demodulation_mode);
    if (int rc = c->use()) return rc;
    rfa_scheduler *s = new rfa_scheduler();
    s->ctx = c;
    s->d = *d;
    s->bps = d->format == RFA_FMT_S16LE ? 4 : 2;
    int rc = RFA_OK;
    do {
        rfa_spectrum_desc sd{};
        sd.format = d->format;
        sd.fft_size = d->fft_size;
        sd.window = d->window;
        sd.avg_len = d->avg_len;
        sd.peak_hold = d->peak_hold;
        if ((rc = rfa_spectrum_plan_create(c, &sd, &s->plan))) break;
        const size_t n = (size_t)d->fft_size;
        if ((rc = s->ring.ensure((size_t)d->ring_rows * n * sizeof(float)))) break;
        if ((rc = s->peaks.ensure(n * sizeof(float)))) break;
        if ((rc = s->avg.ensure(n * sizeof(float)))) break;
        if ((rc = s->partial.ensure(n * s->bps))) break;
        if ((rc = rfa_fill(c, s->ring.as<float>(), (long long)d->ring_rows * (long long)n, -9999.0f))) break;  // FftProcessor.kt:181
        if ((rc = rfa_fill(c, s->peaks.as<float>(), (long long)n, -999999.0f))) break;                          // :236
        if ((rc = rfa_fill(c, s->avg.as<float>(), (long long)n, -9999.0f))) break;
        if (cudaEventCreateWithFlags(&s->ev, cudaEventDisableTiming) != cudaSuccess) {
            rc = RFA_ERR_CUDA;
            break;
        }
        if (d->demodulation_mode != RFA_MODE_OFF) {
            if ((rc = rfa_ctx_create(c->device, nullptr, &s->ctx2))) break;
            s->ctx2->tune = c->tune;
            rfa_chain_desc cd{};
            cd.format = d->format;
            cd.sample_rate = d->sample_rate;
            cd.source_frequency = d->source_frequency;
            cd.channel_frequency = d->channel_frequency;
            cd.mode = d->demodulation_mode;
            cd.channel_width = d->channel_width;
            cd.packet_samples = d->packet_samples;
            cd.volume = d->volume;
            cd.flags = d->flags;
            if ((rc = rfa_chain_create(s->ctx2, &cd, &s->chain))) break;
            if ((rc = rfa_chain_info(s->chain, nullptr, nullptr, nullptr, nullptr, &s->channel_width, nullptr, nullptr))) break;
        }
        s->squelch_satisfied = !d->squelch_enabled;  // AppStateRepository.kt:322: always satisfied when the squelch is disabled
    } while (0);
    if (rc) {
        rfa_scheduler_destroy(s);
        return rc;
    }
    *out = s;
    return RFA_OK;
}

int rfa_scheduler_state(const rfa_scheduler *s, float **ring, long long *newest_row, long long *valid_rows, float **peaks,
                        float **avg, int *squelch_satisfied, int *debounce_counter, long long *packets, long long *frames) {
    RFA_REQUIRE(s != nullptr, "rfa_scheduler_state: NULL");
    if (ring) *ring = s->ring.as<float>();
    if (newest_row) *newest_row = s->history ? (s->write_index + 1) % s->d.ring_rows : -1;
    if (valid_rows) *valid_rows = s->history;
    if (peaks) *peaks = s->peaks.as<float>();
    if (avg) *avg = s->avg.as<float>();
    if (squelch_satisfied) *squelch_satisfied = s->squelch_satisfied ? 1 : 0;
    if (debounce_counter) *debounce_counter = s->debounce;
    if (packets) *packets = s->packets_total;
    if (frames) *frames = s->frames_total;
    return RFA_OK;
}

int rfa_scheduler_read(rfa_scheduler *s, float *ring, float *peaks, float *avg) {
    RFA_REQUIRE(s != nullptr, "rfa_scheduler_read: NULL");
    rfa_ctx *c = s->ctx;
    if (int rc = c->use()) return rc;
    const size_t n = (size_t)s->d.fft_size;
    if (ring) RFA_CK(cudaMemcpyAsync(ring, s->ring.p, (size_t)s->d.ring_rows * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (peaks) RFA_CK(cudaMemcpyAsync(peaks, s->peaks.p, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (avg) RFA_CK(cudaMemcpyAsync(avg, s->avg.p, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    RFA_CK(cudaStreamSynchronize(c->stream));
    return RFA_OK;
}

}  // extern "C"

// one piece of a call: at most ring_rows frames complete in it, so every new row is still in the ring when its
// channel strength is reduced
static int process_piece(rfa_scheduler *s, const void *packets, long long npackets, rfa_scheduler_io *io, int mem) {
    io->frames = 0;
    io->n_audio = 0;
    rfa_ctx *c = s->ctx;
    if (int rc = c->use()) return rc;
    const long long P = s->d.packet_samples, N = s->d.fft_size;
    const int bps = s->bps;
    const size_t packet_bytes = (size_t)P * bps;
    // ---- the packets on the device ------------------------------------------------------------------------------
    const unsigned char *dpk = (const unsigned char *)packets;
    if (mem == RFA_MEM_HOST) {
        if (int rc = s->stage.ensure((size_t)npackets * packet_bytes)) return rc;
        RFA_CK(cudaMemcpyAsync(s->stage.p, packets, (size_t)npackets * packet_bytes, cudaMemcpyHostToDevice, c->stream));
        dpk = s->stage.as<unsigned char>();
    }
    // ---- 1. which samples of which packet make up which frame (Scheduler.kt:254-276) ----------------------------
    std::vector<GatherSeg> segs;
    std::vector<long long> frame_done_at;  // packet index that completed frame j
    long long fill = s->fill, nfr = 0;
    const bool direct = P == N && fill == 0;  // every packet IS a frame: no copy at all
    if (!direct) {
        segs.reserve((size_t)npackets + 1);
        if (fill) segs.push_back(GatherSeg{-1, 0, (int)(fill * bps), 0});  // the carried partial frame comes first
    }
    for (long long k = 0; k < npackets; k++) {
        const long long take = P < N - fill ? P : N - fill;  // fillPacketIntoSamplePacket stops at the buffer's capacity
        if (!direct) segs.push_back(GatherSeg{(long long)(k * packet_bytes), (long long)((nfr * N + fill) * bps), (int)(take * bps), 0});
        fill += take;
        if (fill == N) {
            frame_done_at.push_back(k);
            nfr++;
            fill = 0;
        }
    }
    const unsigned char *dframes = dpk;
    if (!direct) {
        if (int rc = s->frames.ensure((size_t)(nfr + 1) * N * bps)) return rc;
        if (int rc = s->segs.ensure(segs.size() * sizeof(GatherSeg))) return rc;
        unsigned char *fb = s->frames.as<unsigned char>();
        size_t first = 0;
        if (s->fill) {  // carried samples: device-to-device, ahead of the gather
            RFA_CK(cudaMemcpyAsync(fb, s->partial.p, (size_t)s->fill * bps, cudaMemcpyDeviceToDevice, c->stream));
            first = 1;
        }
        const int nseg = (int)(segs.size() - first);
        if (nseg > 0) {
            RFA_CK(cudaMemcpyAsync(s->segs.p, segs.data() + first, (size_t)nseg * sizeof(GatherSeg), cudaMemcpyHostToDevice, c->stream));
            const int bx = (int)((segs[first].bytes / 16 + 255) / 256) > 0 ? (int)((segs[first].bytes / 16 + 255) / 256) : 1;
            dim3 grid((unsigned)(bx > 8 ? 8 : bx), (unsigned)(nseg < 32768 ? nseg : 32768));
            gather_kernel<<<grid, 256, 0, c->stream>>>(dpk, fb, s->segs.as<GatherSeg>(), nseg);
            RFA_CK(cudaGetLastError());
            c->launches++;
        }
        if (fill)  // the new partial frame, kept for the next call
            RFA_CK(cudaMemcpyAsync(s->partial.p, fb + (size_t)nfr * N * bps, (size_t)fill * bps, cudaMemcpyDeviceToDevice, c->stream));
        dframes = fb;
    }
    // ---- gates that do not need the spectrum: chain on its own stream beside the spectrum launch ------------------
    const bool demod = s->chain != nullptr;
    const bool gates_known = !s->d.squelch_enabled;
    std::vector<unsigned char> dem((size_t)npackets, 0), rec((size_t)npackets, 0);
    auto run_chain = [&](const std::vector<unsigned char> &gate) -> int {
        if (!demod) return RFA_OK;
        long long k = 0;
        while (k < npackets) {
            if (!gate[k]) {
                k++;
                continue;
            }
            long long e = k;
            while (e < npackets && gate[e]) e++;
            const long long ns = (e - k) * P;
            float *dst = io->audio ? io->audio + io->n_audio : nullptr;
            RFA_REQUIRE(dst != nullptr, "the demodulator is active: io->audio must be given");
            long long got = 0;
            const void *src = (mem == RFA_MEM_HOST ? (const unsigned char *)packets : dpk) + (size_t)k * packet_bytes;
            if (int rc = rfa_chain_process(s->chain, src, ns, dst, io->audio_capacity - io->n_audio, &got, mem)) return rc;
            io->n_audio += got;
            k = e;
        }
        return RFA_OK;
    };
    // ---- 2. spectrum of every new frame into the reference's ring ------------------------------------------------
    if (nfr > 0) {
        rfa_spectrum_out o{};
        o.rows = s->ring.as<float>();
        o.row0 = s->write_index;
        o.row_step = -1;  // FftProcessor.kt:224-229: the ring is written backwards
        o.ring_rows = s->d.ring_rows;
        o.row_stride = N;
        o.history_rows = s->history;
        o.peaks = s->d.peak_hold ? s->peaks.as<float>() : nullptr;
        o.peaks_accumulate = s->have_peaks ? 1 : 0;
        o.avg = s->avg.as<float>();
        if (int rc = rfa_spectrum_process(s->plan, dframes, nfr, &o, RFA_MEM_DEVICE)) return rc;
        s->have_peaks = true;
    }
    if (gates_known) {
        // Scheduler.kt:161-165 with squelchSatisfied constantly true: the counter stays 0, every gate is open
        for (long long k = 0; k < npackets; k++) {
            dem[k] = demod ? 1 : 0;
            rec[k] = 1;
        }
        s->debounce = 0;
        if (demod && mem == RFA_MEM_DEVICE) {  // the chain's stream must see the caller's packets
            RFA_CK(cudaEventRecord(s->ev, c->stream));
            RFA_CK(cudaStreamWaitEvent(s->ctx2->stream, s->ev, 0));
        }
        if (int rc = run_chain(dem)) return rc;
    }
    // ---- 3. channel strength of every new row (FftProcessor.kt:143-157) ------------------------------------------
    std::vector<float> strength((size_t)nfr, -999.0f);
    bool strength_valid = false;
    if (nfr > 0 && demod) {  // getChannelFrequencyRange is null without a demodulator: the strength is never updated
        int b0 = 0, b1 = 0;
        const long long cs = s->d.channel_frequency - s->channel_width, ce = s->d.channel_frequency + s->channel_width;
        if (int rc = rfa_channel_bins((int)N, s->d.source_frequency, s->d.sample_rate, cs, ce, &b0, &b1)) return rc;
        if (b1 > b0) {
            if (int rc = s->strength.ensure((size_t)nfr * sizeof(float))) return rc;
            // rows of the new frames: write_index, write_index - 1, ... (mod ring); frames older than the ring were never stored
            const long long stored = nfr < s->d.ring_rows ? nfr : s->d.ring_rows;
            const long long first_stored = nfr - stored;
            long long row0 = (s->write_index - first_stored) % s->d.ring_rows;
            if (row0 < 0) row0 += s->d.ring_rows;
            if (int rc = rfa_channel_strength(c, s->ring.as<float>(), row0, -1, s->d.ring_rows, N, stored, b0, b1,
                                              s->strength.as<float>() + first_stored, RFA_MEM_DEVICE))
                return rc;
            RFA_CK(cudaMemcpyAsync(strength.data() + first_stored, s->strength.as<float>() + first_stored, (size_t)stored * sizeof(float),
                                   cudaMemcpyDeviceToHost, c->stream));
            RFA_CK(cudaStreamSynchronize(c->stream));
            RFA_REQUIRE(first_stored == 0, "internal: a piece completed more frames than the ring holds");
            strength_valid = true;
        }
    }
    if (nfr > 0) {
        s->write_index = ((s->write_index - nfr) % s->d.ring_rows + s->d.ring_rows) % s->d.ring_rows;
        s->history = s->history + nfr > s->d.ring_rows ? s->d.ring_rows : s->history + nfr;
    }
    // ---- 4. the gate sequence, packet by packet (Scheduler.kt:161-165, :199, :237) --------------------------------
    if (!gates_known) {
        size_t next_frame = 0;
        for (long long k = 0; k < npackets; k++) {
            if (s->squelch_satisfied)
                s->debounce = 0;
            else if (s->debounce < kSquelchDebounceCount)
                s->debounce++;
            const bool open = s->squelch_satisfied || s->debounce < kSquelchDebounceCount;
            rec[k] = (s->squelch_satisfied || !s->d.record_only_when_squelch_satisfied || s->debounce < kSquelchDebounceCount) ? 1 : 0;
            dem[k] = (demod && open) ? 1 : 0;
            // the frame this packet completed reaches the FftProcessor, which updates the strength before the next packet
            if (next_frame < frame_done_at.size() && frame_done_at[next_frame] == k) {
                if (strength_valid) s->squelch_satisfied = strength[next_frame] > s->d.squelch_db;  // AppStateRepository.kt:320-321
                next_frame++;
            }
        }
        // ---- 5. the delivered packets through the chain ------------------------------------------------------------
        if (demod && mem == RFA_MEM_DEVICE) {
            RFA_CK(cudaEventRecord(s->ev, c->stream));
            RFA_CK(cudaStreamWaitEvent(s->ctx2->stream, s->ev, 0));
        }
        if (int rc = run_chain(dem)) return rc;
    }
    if (demod && mem == RFA_MEM_DEVICE) {
        // device-buffer chain calls only enqueue work on the chain's own stream: the call returns with the audio written
        RFA_CK(cudaEventRecord(s->ev, s->ctx2->stream));
        RFA_CK(cudaStreamWaitEvent(c->stream, s->ev, 0));
    }
    RFA_CK(cudaStreamSynchronize(c->stream));
    s->fill = fill;
    s->packets_total += npackets;
    s->frames_total += nfr;
    io->frames = nfr;
    if (io->signal_strength)
        for (long long j = 0; j < nfr; j++) io->signal_strength[j] = strength[j];
    if (io->demod_gate) memcpy(io->demod_gate, dem.data(), (size_t)npackets);
    if (io->record_gate) memcpy(io->record_gate, rec.data(), (size_t)npackets);
    io->packets = npackets;
    return RFA_OK;
}

extern "C" int rfa_scheduler_process(rfa_scheduler *s, const void *packets, long long npackets, rfa_scheduler_io *io, int mem) {
    RFA_REQUIRE(s && io, "rfa_scheduler_process: NULL argument");
    io->frames = 0;
    io->n_audio = 0;
    io->packets = 0;
    RFA_REQUIRE(npackets >= 0, "negative packet count");
    if (npackets == 0) return RFA_OK;
    RFA_REQUIRE(packets != nullptr, "packets is NULL");
    RFA_REQUIRE(npackets < (1LL << 24), "too many packets in one call");
    const long long P = s->d.packet_samples, N = s->d.fft_size;
    const size_t packet_bytes = (size_t)P * s->bps;
    long long done = 0;
    while (done < npackets) {
        // packets of this piece: up to the one that completes frame number ring_rows
        long long fill = s->fill, frames = 0, k = done;
        for (; k < npackets && frames < s->d.ring_rows; k++) {
            fill += P < N - fill ? P : N - fill;
            if (fill == N) {
                frames++;
                fill = 0;
            }
        }
        rfa_scheduler_io piece = *io;
        if (io->signal_strength) piece.signal_strength = io->signal_strength + io->frames;
        if (io->demod_gate) piece.demod_gate = io->demod_gate + done;
        if (io->record_gate) piece.record_gate = io->record_gate + done;
        if (io->audio) piece.audio = io->audio + io->n_audio;
        piece.audio_capacity = io->audio_capacity - io->n_audio;
        if (int rc = process_piece(s, (const unsigned char *)packets + (size_t)done * packet_bytes, k - done, &piece, mem)) return rc;
        io->frames += piece.frames;
        io->n_audio += piece.n_audio;
        done = k;
    }
    io->packets = npackets;
    return RFA_OK;
}
