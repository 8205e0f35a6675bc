// recording.cu -- IQ recordings on disk: the reference's file format, its
// file-name metadata convention, its FileIQSource packet semantics, and a streamed spectrum pass
// that reads the file through pinned double buffers while the GPU transforms the previous chunk.
//
//   IQ_FILE_FORMAT.md:1-120            header-less interleaved IQ, HACKRF / RTLSDR / AIRSPY / HYDRASDR
//   A/database/RecordingDao.kt:87-90   Recording.calculateFileName()
//   A/ui/composable/HelperComposables.kt:168-179  Long.asStringWithUnit()
//   A/ui/MainViewModel.kt:2034-2080    setFilesourceUri(): metadata from the file name
//   A/source/FileIQSource.java:305-369 getPacket(): whole packets only, rewind on repeat, real-time pacing
#include <ctype.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>

#include "../../include/rfa_b200.h"
#include "capi_core.h"

using namespace rfa;

namespace {

// does `s` contain, anywhere, a separator (_ - whitespace) followed by digits followed directly by one of
// the unit spellings?  Mirrors  filename.matches(".*(_|-|\\s)([0-9]+)(u1|u2|..).*")  with the greedy
// leading ".*": the LAST possible match wins, group 2 is the whole digit run after the separator.
bool find_number_with_unit(const std::string &s, const char *const *units, int nunits, long long *value) {
    bool found = false;
    for (size_t i = 0; i + 1 < s.size(); i++) {
        const char c = s[i];
        if (!(c == '_' || c == '-' || isspace((unsigned char)c))) continue;
        size_t j = i + 1;
        while (j < s.size() && isdigit((unsigned char)s[j])) j++;
        if (j == i + 1) continue;
        for (int u = 0; u < nunits; u++) {
            const size_t len = strlen(units[u]);
            if (s.compare(j, len, units[u]) == 0) {
                // Kotlin's toLong(): more than 18 digits overflow -> NumberFormatException -> value untouched
                if (j - (i + 1) > 18) return found;
                *value = atoll(s.substr(i + 1, j - (i + 1)).c_str());
                found = true;
                break;
            }
        }
    }
    return found;
}

bool contains_any(const std::string &s, const char *const *words, int n) {
    for (int i = 0; i < n; i++)
        if (s.find(words[i]) != std::string::npos) return true;
    return false;
}

}  // namespace

extern "C" {

int rfa_recording_parse_name(const char *filename, rfa_recording_info *info) {
    RFA_REQUIRE(filename && info, "rfa_recording_parse_name: NULL argument");
    const std::string f(filename);
    // 1. format (MainViewModel.kt:2042-2054; later matches override earlier ones)
    static const char *const hackrf[] = {"hackrf", "HackRF", "HACKRF", "hackrfone"};
    static const char *const rtlsdr[] = {"rtlsdr", "rtl-sdr", "RTLSDR", "RTL-SDR"};
    static const char *const airspy[] = {"airspy", "Airspy", "AIRSPY", "AirSpy"};
    static const char *const hydra[] = {"hydrasdr", "HydraSDR", "HYDRASDR", "HydraSdr"};
    if (contains_any(f, hackrf, 4)) info->file_format = RFA_FILE_HACKRF, info->have_format = 1;
    if (contains_any(f, rtlsdr, 4)) info->file_format = RFA_FILE_RTLSDR, info->have_format = 1;
    if (contains_any(f, airspy, 4)) info->file_format = RFA_FILE_AIRSPY, info->have_format = 1;
    if (contains_any(f, hydra, 4)) info->file_format = RFA_FILE_HYDRASDR, info->have_format = 1;
    // 2. sample rate (:2056-2064), 3. frequency (:2066-2074): plain, kilo, mega -- in that order
    static const char *const sps[] = {"sps", "Sps", "SPS"}, *const ksps[] = {"ksps", "Ksps", "KSps", "KSPS"},
                      *const msps[] = {"msps", "Msps", "MSps", "MSPS"};
    static const char *const hz[] = {"hz", "Hz", "HZ"}, *const khz[] = {"khz", "Khz", "KHz", "KHZ"},
                      *const mhz[] = {"mhz", "Mhz", "MHz", "MHZ"};
    long long v;
    if (find_number_with_unit(f, sps, 3, &v)) info->sample_rate = v, info->have_sample_rate = 1;
    if (find_number_with_unit(f, ksps, 4, &v)) info->sample_rate = v * 1000, info->have_sample_rate = 1;
    if (find_number_with_unit(f, msps, 4, &v)) info->sample_rate = v * 1000000, info->have_sample_rate = 1;
    if (find_number_with_unit(f, hz, 3, &v)) info->frequency = v, info->have_frequency = 1;
    if (find_number_with_unit(f, khz, 4, &v)) info->frequency = v * 1000, info->have_frequency = 1;
    if (find_number_with_unit(f, mhz, 4, &v)) info->frequency = v * 1000000, info->have_frequency = 1;
    return RFA_OK;
}

int rfa_recording_sample_format(int file_format) {  // IQ_FILE_FORMAT.md:26-84
    switch (file_format) {
        case RFA_FILE_HACKRF: return RFA_FMT_S8;
        case RFA_FILE_RTLSDR: return RFA_FMT_U8;
        case RFA_FILE_AIRSPY:
        case RFA_FILE_HYDRASDR: return RFA_FMT_S16LE;
    }
    return -1;
}

// Long.asStringWithUnit(unit).replace(" ", "")  (HelperComposables.kt:168-179, RecordingDao.kt:89)
static std::string with_unit(long long value, const char *unit) {
    static const char *const prefix[] = {"", "k", "M", "G", "T"};
    int index = 0;
    while (value % 1000 == 0 && value >= 1000 && index < 4) {
        value /= 1000;
        index++;
    }
    return std::to_string(value) + prefix[index] + unit;  // the grouping separator is a space, removed again
}

int rfa_recording_file_name(const char *timestamp, const char *name, int file_format, long long frequency,
                            long long sample_rate, char *out, int capacity) {
    RFA_REQUIRE(timestamp && name && out && capacity > 0, "rfa_recording_file_name: NULL argument");
    static const char *const fmt[] = {"HACKRF", "RTLSDR", "AIRSPY", "HYDRASDR"};
    RFA_REQUIRE(file_format >= RFA_FILE_HACKRF && file_format <= RFA_FILE_HYDRASDR, "unknown file format %d", file_format);
    const std::string s = std::string(timestamp) + "_" + name + "_" + fmt[file_format] + "_" + with_unit(frequency, "Hz") +
                          "_" + with_unit(sample_rate, "Sps") + ".iq";
    RFA_REQUIRE((int)s.size() < capacity, "file name needs %zu bytes", s.size() + 1);
    memcpy(out, s.c_str(), s.size() + 1);
    return RFA_OK;
}

}  // extern "C"

// ---- FileIQSource.getPacket ---------------------------------------------------------------------
struct rfa_file_source {
    std::string path;
    FILE *fp = nullptr;
    long long packet_bytes = 0;
    int bytes_per_sample = 2;
    bool repeat = false;
    long long sample_rate = 0;  // > 0: pace like the hardware would (FileIQSource.java:343-353)
    long long bytes_read = 0;
    std::chrono::steady_clock::time_point start;
};

extern "C" {

int rfa_file_source_open(const char *path, int file_format, long long packet_bytes, int repeat,
                         long long pace_sample_rate, rfa_file_source **out) {
    RFA_REQUIRE(path && out, "rfa_file_source_open: NULL argument");
    *out = nullptr;
    const int sf = rfa_recording_sample_format(file_format);
    RFA_REQUIRE(sf >= 0, "Invalid file format: %d", file_format);  // FileIQSource.java:86
    RFA_REQUIRE(packet_bytes > 0, "packet size must be positive");
    FILE *fp = fopen(path, "rb");
    if (!fp) {
        set_error("Error while opening file: %s", path);  // FileIQSource.java:113
        return RFA_ERR_INVALID;
    }
    rfa_file_source *s = new rfa_file_source();
    s->path = path;
    s->fp = fp;
    s->packet_bytes = packet_bytes;
    s->bytes_per_sample = sf == RFA_FMT_S16LE ? 4 : 2;
    s->repeat = repeat != 0;
    s->sample_rate = pace_sample_rate;
    s->start = std::chrono::steady_clock::now();  // startSampling()
    *out = s;
    return RFA_OK;
}

int rfa_file_source_close(rfa_file_source *s) {
    if (!s) return RFA_OK;
    if (s->fp) fclose(s->fp);
    delete s;
    return RFA_OK;
}

// 1 = a whole packet was written to `packet`, 0 = end of file ("End of File", a trailing partial packet is
// dropped like the reference drops it), < 0 = error.
int rfa_file_source_get_packet(rfa_file_source *s, void *packet) {
    if (!s || !packet || !s->fp) return -RFA_ERR_INVALID;
    int got = 0;
    if ((long long)fread(packet, 1, (size_t)s->packet_bytes, s->fp) == s->packet_bytes) {
        got = 1;
    } else if (s->repeat) {  // rewind and try again (:326-338)
        fclose(s->fp);
        s->fp = fopen(s->path.c_str(), "rb");
        if (!s->fp) {
            set_error("Error while re-openening file");
            return -RFA_ERR_INVALID;
        }
        if ((long long)fread(packet, 1, (size_t)s->packet_bytes, s->fp) == s->packet_bytes) got = 1;
    }
    if (got) s->bytes_read += s->packet_bytes;
    if (s->sample_rate > 0) {  // simulate the sample rate of real hardware
        const double ns_per_sample = 1e9 / (double)s->sample_rate;
        const auto expected = s->start + std::chrono::nanoseconds((long long)(ns_per_sample * (double)s->bytes_read / s->bytes_per_sample));
        std::this_thread::sleep_until(expected);
    }
    return got;
}

long long rfa_file_source_bytes_read(const rfa_file_source *s) { return s ? s->bytes_read : 0; }

// ---- streamed spectrum pass over a recording -------------------------------------------------------------
// Frames [first_frame, first_frame + nframes) of the file (nframes < 0: to the end; a trailing partial frame is
// dropped).  A reader thread fills two pinned buffers alternately; the calling thread hands each filled
// buffer to rfa_spectrum_process (host mode: chunked H2D -> kernel -> D2H on three streams).  out->rows,
// when given, is a host array of all rows; peaks accumulate over the file; avg is that of the file's newest rows.
int rfa_spectrum_process_file(rfa_spectrum_plan *plan, const char *path, long long first_frame, long long nframes,
                              const rfa_spectrum_out *out, long long chunk_frames, long long *frames_done) {
    RFA_REQUIRE(plan && path && out, "rfa_spectrum_process_file: NULL argument");
    if (frames_done) *frames_done = 0;
    int n = 0, fmt = 0, L = 0;
    if (int rc = rfa_spectrum_plan_info(plan, &n, &fmt, &L)) return rc;
    const long long frame_bytes = (long long)n * (fmt == RFA_FMT_S16LE ? 4 : 2);
    FILE *fp = fopen(path, "rb");
    if (!fp) {
        set_error("Error while opening file: %s", path);
        return RFA_ERR_INVALID;
    }
    fseeko(fp, 0, SEEK_END);
    const long long file_frames = (long long)ftello(fp) / frame_bytes;
    if (first_frame < 0 || first_frame > file_frames) {
        fclose(fp);
        set_error("first frame %lld outside the file (%lld frames)", first_frame, file_frames);
        return RFA_ERR_INVALID;
    }
    long long total = file_frames - first_frame;
    if (nframes >= 0 && nframes < total) total = nframes;
    if (total == 0) {
        fclose(fp);
        return RFA_OK;
    }
    if (chunk_frames <= 0) chunk_frames = (32LL << 20) / frame_bytes;  // 32 MiB of IQ per chunk
    if (chunk_frames < L + 1) chunk_frames = L + 1;
    // a short last chunk is merged into its predecessor so that the final average sees L+1 rows of one call
    long long nchunks = total / chunk_frames;
    if (nchunks == 0) nchunks = 1;
    const long long max_chunk = total < chunk_frames ? total : chunk_frames + total % chunk_frames;
    Buf pin[2];
    pin[0].pinned = pin[1].pinned = true;
    const size_t cap = (size_t)max_chunk * frame_bytes;
    int rc = pin[0].ensure(cap);
    if (!rc) rc = pin[1].ensure(cap);
    if (rc) {
        fclose(fp);
        pin[0].release();
        pin[1].release();
        return rc;
    }
    fseeko(fp, (off_t)(first_frame * frame_bytes), SEEK_SET);
    auto frames_of = [&](long long i) { return i == nchunks - 1 ? total - i * chunk_frames : chunk_frames; };
    // reader thread: chunk i goes to buffer i & 1 once the consumer has released it
    std::mutex mu;
    std::condition_variable cv;
    long long filled = 0, released = 0;  // chunks read / chunks the GPU side is done with
    bool read_error = false;
    std::thread reader([&] {
        for (long long i = 0; i < nchunks; i++) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return i - released < 2; });
            }
            const size_t want = (size_t)(frames_of(i) * frame_bytes);
            const bool ok = fread(pin[i & 1].p, 1, want, fp) == want;
            std::lock_guard<std::mutex> lk(mu);
            if (!ok) read_error = true;
            filled = i + 1;
            cv.notify_all();
            if (!ok) return;
        }
    });
    long long done = 0;
    for (long long i = 0; i < nchunks && !rc; i++) {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return filled > i; });
            if (read_error && filled == i + 1) {  // the chunk that failed is the one we were waiting for
                set_error("Unexpected error while reading file: %s", path);  // FileIQSource.java:356
                rc = RFA_ERR_INVALID;
            }
        }
        if (rc) break;
        const long long frames = frames_of(i);
        rfa_spectrum_out o = *out;
        if (out->rows) o.rows = out->rows + (size_t)done * (size_t)(out->row_stride > 0 ? out->row_stride : n);
        o.row0 = 0;
        o.row_step = 1;
        o.ring_rows = 0;
        o.history_rows = 0;
        o.peaks_accumulate = (out->peaks_accumulate || i > 0) ? 1 : 0;
        o.avg = (i == nchunks - 1) ? out->avg : nullptr;
        rc = rfa_spectrum_process(plan, pin[i & 1].p, frames, &o, RFA_MEM_HOST);
        done += frames;
        std::lock_guard<std::mutex> lk(mu);
        released = i + 1;
        cv.notify_all();
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        released = nchunks + 2;  // let a waiting reader run to its end
        cv.notify_all();
    }
    reader.join();
    fclose(fp);
    pin[0].release();
    pin[1].release();
    if (frames_done) *frames_done = done;
    return rc;
}

}  // extern "C"
