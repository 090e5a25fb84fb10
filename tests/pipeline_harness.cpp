// CPU-only harness for the product's ordered multi-handle pipeline (cedarx_h264_encoder_b200/csrc/pipeline.cpp), which
// uses nothing but the public C ABI: this file supplies a stand-in for the seven encoder calls it makes (a "coder" that
// turns every frame into a record of its stream position, a hash of its bytes and its picture type, with random delays
// to shake the thread interleaving) and drives the pipe_* calls through the cases tests/test_pipeline_host.py lists.
// Built with ThreadSanitizer and with AddressSanitizer by that test.  TEST INFRASTRUCTURE: nothing here is shipped.
//
// What is checked is what the concatenation must honour (kernel/cedar.c:1047-1061, 1193-1196): batches come back in
// submission order, every batch is told its position in the stream (first_frame_index), only the last one may be short.
#include "../include/cedar_b200.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <errno.h>
#include <random>
#include <thread>
#include <vector>

// ---------------------------------------------------------------------------------------------------------------------
// stand-in encoder
// ---------------------------------------------------------------------------------------------------------------------
struct cedar_b200_handle {
    cedar_b200_config cfg;
    size_t fb;
    std::vector<uint8_t> staging, dev, out;
    std::vector<int> sizes;
    std::vector<double> sse;
    int n = 0;
    std::mt19937 rng;
};

static std::atomic<int> g_open_handles{0}, g_fail_at_frame{-1};

static uint32_t fnv(const uint8_t *p, size_t n)
{
    uint32_t h = 2166136261u;
    for (size_t i = 0; i < n; i++)
        h = (h ^ p[i]) * 16777619u;
    return h;
}

static void nap(cedar_b200_handle *h, int max_us)
{
    std::this_thread::sleep_for(std::chrono::microseconds(h->rng() % (unsigned)(max_us + 1)));
}

extern "C" {
int cedar_b200_open(const struct cedar_b200_config *cfg, struct cedar_b200_io *io, cedar_b200_handle **out)
{
    if (cfg->qp == 99)
        return -ENODEV; // open failure injected by the test
    cedar_b200_handle *h = new cedar_b200_handle();
    h->cfg = *cfg;
    h->fb = (size_t)cfg->src_width * cfg->src_height * 3 / 2;
    h->staging.resize(h->fb * (size_t)cfg->max_clip_frames);
    h->dev.resize(h->staging.size());
    h->sizes.resize((size_t)cfg->max_clip_frames);
    h->sse.resize((size_t)cfg->max_clip_frames);
    h->rng.seed((unsigned)(1234 + cfg->device * 77 + g_open_handles.fetch_add(1)));
    memset(io, 0, sizeof(*io));
    *out = h;
    return 0;
}
void cedar_b200_close(cedar_b200_handle *h)
{
    g_open_handles.fetch_sub(1);
    delete h;
}
void *cedar_b200_clip_input(cedar_b200_handle *h, size_t *frame_bytes)
{
    if (frame_bytes)
        *frame_bytes = h->fb;
    return h->staging.data();
}
int cedar_b200_clip_upload(cedar_b200_handle *h, int nframes)
{
    if (nframes <= 0 || nframes > h->cfg.max_clip_frames)
        return -EINVAL;
    nap(h, 300);
    memcpy(h->dev.data(), h->staging.data(), h->fb * (size_t)nframes);
    return 0;
}
int cedar_b200_clip_encode(cedar_b200_handle *h, int nframes, int first_frame_index)
{
    if (nframes <= 0 || nframes > h->cfg.max_clip_frames || first_frame_index % h->cfg.keyframe_interval)
        return -EINVAL;
    nap(h, 2000);
    const int fail = g_fail_at_frame.load();
    if (fail >= first_frame_index && fail < first_frame_index + nframes)
        return -EIO;
    h->out.clear();
    for (int i = 0; i < nframes; i++) {
        const uint32_t f = (uint32_t)(first_frame_index + i), hash = fnv(h->dev.data() + h->fb * (size_t)i, h->fb);
        const size_t at = h->out.size(), sz = 10 + f % 5;
        h->out.resize(at + sz, 0xee);
        memcpy(&h->out[at], &f, 4);
        memcpy(&h->out[at + 4], &hash, 4);
        h->out[at + 8] = (uint8_t)(f % (uint32_t)h->cfg.keyframe_interval == 0); // IDR
        h->out[at + 9] = (uint8_t)(f == 0);                                      // carries SPS + PPS
        h->sizes[(size_t)i] = (int)sz;
        h->sse[(size_t)i] = (double)f;
    }
    h->n = nframes;
    return 0;
}
long long cedar_b200_clip_download(cedar_b200_handle *h, const uint8_t **out, int *frame_bytes)
{
    nap(h, 500);
    *out = h->out.data();
    memcpy(frame_bytes, h->sizes.data(), sizeof(int) * (size_t)h->n);
    return (long long)h->out.size();
}
int cedar_b200_stats(cedar_b200_handle *h, double *sse_y, int nframes)
{
    memcpy(sse_y, h->sse.data(), sizeof(double) * (size_t)nframes);
    return 0;
}
}

// ---------------------------------------------------------------------------------------------------------------------
// cases
// ---------------------------------------------------------------------------------------------------------------------
#define CHECK(c)                                                                                                          \
    do {                                                                                                                  \
        if (!(c)) {                                                                                                       \
            fprintf(stderr, "pipeline_harness: %s:%d: check failed: %s\n", __FILE__, __LINE__, #c);                       \
            exit(1);                                                                                                      \
        }                                                                                                                 \
    } while (0)

static const int W_ = 16, H_ = 8; // 192-byte frames

static void fill_frame(uint8_t *dst, size_t fb, int f)
{
    for (size_t i = 0; i < fb; i++)
        dst[i] = (uint8_t)(f * 31 + (int)i * 7 + (f >> 3));
}

static cedar_b200_config make_cfg(int gop, int qp = 24)
{
    cedar_b200_config c;
    memset(&c, 0, sizeof(c));
    c.src_width = c.dst_width = W_;
    c.src_height = c.dst_height = H_;
    c.qp = qp;
    c.keyframe_interval = gop;
    return c;
}

// One stream of `total` frames through `handles` x `ndev` workers; `end` = how the producer ends it when the last batch
// is full (0: finish(), 1: submit(0)); `poll`: the consumer polls with wait = 0; `release`: it calls pipe_release.
static void run_stream(int ndev, int handles, int gop, int batch_gops, int total, int end, bool poll, bool release, int fail_at)
{
    cedar_b200_config cfg = make_cfg(gop);
    int devs[8];
    for (int i = 0; i < ndev; i++)
        devs[i] = i;
    cedar_b200_pipe *p = nullptr;
    g_fail_at_frame = fail_at;
    CHECK(cedar_b200_pipe_open(&cfg, devs, ndev, handles, batch_gops, &p) == 0 && p);
    CHECK(cedar_b200_pipe_workers(p) == ndev * (handles ? handles : 2));
    std::thread producer([&] {
        int f = 0;
        for (;;) {
            size_t fb = 0;
            int cap = 0, n = 0;
            uint8_t *in = (uint8_t *)cedar_b200_pipe_acquire(p, &fb, &cap);
            if (!in)
                break;
            CHECK(fb == (size_t)W_ * H_ * 3 / 2 && cap == (batch_gops ? batch_gops : 4) * gop);
            for (; n < cap && f < total; n++, f++)
                fill_frame(in + fb * (size_t)n, fb, f);
            if (n == 0 && end == 0) { // nothing left and the producer prefers finish(): the slot still has to go back
                CHECK(cedar_b200_pipe_submit(p, 0) == 0);
                break;
            }
            CHECK(cedar_b200_pipe_submit(p, n) == 0);
            if (n < cap)
                break;
            if (f == total && end == 0)
                break;
        }
        cedar_b200_pipe_finish(p);
    });
    int next_frame = 0, failed_batches = 0;
    const size_t fb = (size_t)W_ * H_ * 3 / 2;
    std::vector<uint8_t> frame(fb);
    for (;;) {
        const uint8_t *out = nullptr;
        const int *sizes = nullptr;
        const double *sse = nullptr;
        int n = 0;
        long long tot = cedar_b200_pipe_next(p, &out, &sizes, &n, &sse, poll ? 0 : 1);
        if (tot == -EAGAIN) {
            std::this_thread::sleep_for(std::chrono::microseconds(200));
            continue;
        }
        if (tot == 0) {
            if (poll && next_frame < total) { // polling: "nothing outstanding" also before the first submit
                std::this_thread::sleep_for(std::chrono::microseconds(200));
                continue;
            }
            break;
        }
        if (tot < 0) { // the injected failure: that batch only
            CHECK(tot == -EIO && fail_at >= next_frame && fail_at < next_frame + n && n > 0);
            next_frame += n;
            failed_batches++;
            continue;
        }
        long long seen = 0;
        for (int i = 0; i < n; i++) {
            uint32_t f, hash;
            memcpy(&f, out + seen, 4);
            memcpy(&hash, out + seen + 4, 4);
            CHECK((int)f == next_frame);                       // submission order, position in the stream
            fill_frame(frame.data(), fb, next_frame);
            CHECK(hash == fnv(frame.data(), fb));              // the frames the producer wrote, nobody else's
            CHECK(out[seen + 8] == (next_frame % gop == 0));   // IDR at every multiple of the keyframe interval
            CHECK(out[seen + 9] == (next_frame == 0));         // parameter sets once
            CHECK(sizes[i] == 10 + next_frame % 5 && sse[i] == (double)next_frame);
            seen += sizes[i];
            next_frame++;
        }
        CHECK(seen == tot);
        if (release)
            CHECK(cedar_b200_pipe_release(p) == 0);
    }
    producer.join();
    CHECK(next_frame == total);
    CHECK(failed_batches == (fail_at >= 0 && fail_at < total ? 1 : 0));
    CHECK(cedar_b200_pipe_next(p, nullptr, nullptr, nullptr, nullptr, 1) == 0); // drained stays drained
    cedar_b200_pipe_close(p);
    CHECK(g_open_handles.load() == 0);
    g_fail_at_frame = -1;
}

static void misuse_and_teardown()
{
    cedar_b200_config cfg = make_cfg(3);
    cedar_b200_pipe *p = nullptr;
    CHECK(cedar_b200_pipe_open(nullptr, nullptr, 0, 0, 0, &p) == -EINVAL);
    CHECK(cedar_b200_pipe_open(&cfg, nullptr, 0, -1, 0, &p) == -EINVAL);
    cedar_b200_config bad = make_cfg(3, 99);
    CHECK(cedar_b200_pipe_open(&bad, nullptr, 0, 3, 1, &p) == -ENODEV && g_open_handles.load() == 0); // every handle closed again
    CHECK(cedar_b200_pipe_open(&cfg, nullptr, 0, 2, 1, &p) == 0);
    CHECK(cedar_b200_pipe_submit(p, 1) == -EINVAL); // nothing acquired
    size_t fb = 0;
    int cap = 0;
    uint8_t *a = (uint8_t *)cedar_b200_pipe_acquire(p, &fb, &cap);
    CHECK(a && cap == 3);
    CHECK(cedar_b200_pipe_acquire(p, nullptr, nullptr) == nullptr); // one batch is filled at a time
    CHECK(cedar_b200_pipe_submit(p, 4) == -EINVAL && cedar_b200_pipe_submit(p, -1) == -EINVAL);
    fill_frame(a, fb, 0), fill_frame(a + fb, fb, 1), fill_frame(a + 2 * fb, fb, 2);
    CHECK(cedar_b200_pipe_submit(p, 3) == 0);
    a = (uint8_t *)cedar_b200_pipe_acquire(p, &fb, &cap);
    CHECK(a);
    fill_frame(a, fb, 3);
    CHECK(cedar_b200_pipe_submit(p, 1) == 0);                       // short: the last batch of the stream
    CHECK(cedar_b200_pipe_acquire(p, nullptr, nullptr) == nullptr); // ... so there is no further one
    CHECK(cedar_b200_pipe_release(p) == 0);                         // nothing handed out yet: a no-op
    // close with both batches submitted and never consumed: waits for the running encodes, then tears down
    cedar_b200_pipe_close(p);
    CHECK(g_open_handles.load() == 0);
    cedar_b200_pipe_close(nullptr);
    CHECK(cedar_b200_pipe_workers(nullptr) == 0 && cedar_b200_pipe_submit(nullptr, 1) == -EINVAL &&
          cedar_b200_pipe_next(nullptr, nullptr, nullptr, nullptr, nullptr, 0) == -EINVAL);
    // finish() before anything was submitted; acquire afterwards
    CHECK(cedar_b200_pipe_open(&cfg, nullptr, 0, 1, 2, &p) == 0);
    CHECK(cedar_b200_pipe_finish(p) == 0);
    CHECK(cedar_b200_pipe_acquire(p, nullptr, nullptr) == nullptr);
    CHECK(cedar_b200_pipe_next(p, nullptr, nullptr, nullptr, nullptr, 1) == 0);
    cedar_b200_pipe_close(p);
    CHECK(g_open_handles.load() == 0);
}

int main(int argc, char **argv)
{
    const int rounds = argc > 1 ? atoi(argv[1]) : 1;
    for (int r = 0; r < rounds; r++) {
        misuse_and_teardown();
        const int shapes[][5] = {// devices, handles, gop, batch_gops, frames
                                 {1, 1, 3, 1, 10}, {1, 2, 3, 2, 20}, {1, 3, 4, 1, 12}, {2, 2, 2, 2, 37}, {4, 1, 5, 1, 50},
                                 {1, 0, 2, 0, 19}, {3, 2, 1, 1, 11}, {1, 2, 3, 2, 0},  {1, 5, 2, 1, 3},  {2, 1, 60, 1, 61}};
        for (const auto &s : shapes)
            for (int variant = 0; variant < 4; variant++)
                run_stream(s[0], s[1], s[2], s[3], s[4], variant & 1, (variant & 2) != 0, variant == 1, -1);
        run_stream(1, 2, 3, 1, 18, 0, false, false, 7);  // one batch fails: reported in its place, the rest unaffected
        run_stream(2, 2, 2, 2, 30, 1, true, true, 0);
        run_stream(1, 1, 4, 1, 9, 0, false, true, 8);
    }
    printf("pipeline_harness ok\n");
    return 0;
}
