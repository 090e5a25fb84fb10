// Ordered multi-handle / multi-GPU pipeline over the clip calls of the C ABI (include/cedar_b200.h).
//
// North star / SURVEY 8e: "GOP-parallel across the GPUs of one box, host thread per GPU, per-GPU bytestreams
// concatenated on the host".  What the concatenation honours is the reference's stream layout: SPS + PPS once, before
// stream frame 0 (kernel/cedar.c:1058-1061); IDR iff frame_p_count == 0 (:1047-1050); the GOP counter restarts at every
// IDR (:1193-1196); idr_pic_id 0 (:1004-1005).  A closed GOP therefore depends on nothing but its own frames and its
// position in the stream, which cedar_b200_clip_encode takes as first_frame_index.
//
// Shape: batches of whole GOPs; batch b goes to worker b % W; a worker is one encoder handle on one device, driven by
// its own host thread (handles share nothing, so the GPU interleaves the chains of several handles by itself, and
// different devices run in parallel).  A producer fills staging buffers (acquire -> submit), a consumer takes finished
// batches in submission order (next), so reading, encoding and writing overlap.  This file uses only the public C ABI.
#include "../../include/cedar_b200.h"

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <errno.h>
#include <mutex>
#include <thread>
#include <vector>

namespace {

enum SlotState { FREE, FILLING, QUEUED, DONE, CONSUMING };

struct Worker {
    cedar_b200_handle *h = nullptr;
    int device = 0;
    std::thread thr;
    // one batch at a time per worker
    SlotState state = FREE;
    int nframes = 0, first_frame = 0;
    long long total = 0; // bytes, or negative errno
    const uint8_t *out = nullptr;
    std::vector<int> sizes;
    std::vector<double> sse;
    uint8_t *staging = nullptr;
};

} // namespace

struct cedar_b200_pipe {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<Worker> w;
    int K = 1, cap = 0; // keyframe interval, frames per batch
    size_t frame_bytes = 0;
    long long submitted = 0, consumed = 0, acquired = 0; // batch counters
    long long frames_submitted = 0;
    bool partial_seen = false, finished = false, stop = false;
    long long consuming = -1; // batch handed out by the last pipe_next
};

namespace {

void worker_main(cedar_b200_pipe *p, int wi)
{
    Worker &w = p->w[wi];
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(p->mu);
            p->cv.wait(lk, [&] { return p->stop || w.state == QUEUED; });
            if (p->stop)
                return;
        }
        static const bool trace = getenv("CEDAR_B200_TRACE") != nullptr;
        auto t0 = std::chrono::steady_clock::now();
        long long total = cedar_b200_clip_upload(w.h, w.nframes);
        if (total == 0)
            total = cedar_b200_clip_encode(w.h, w.nframes, w.first_frame);
        auto t1 = std::chrono::steady_clock::now();
        if (total == 0)
            total = cedar_b200_clip_download(w.h, &w.out, w.sizes.data());
        if (total > 0)
            cedar_b200_stats(w.h, w.sse.data(), w.nframes);
        if (trace)
            fprintf(stderr, "[trace] worker %d (device %d): %d frames from %d: issue %.1f ms, until downloaded %.1f ms\n", wi,
                    w.device, w.nframes, w.first_frame, std::chrono::duration<double, std::milli>(t1 - t0).count(),
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
        {
            std::lock_guard<std::mutex> lk(p->mu);
            w.total = total;
            w.state = DONE;
        }
        p->cv.notify_all();
    }
}

} // namespace

extern "C" {

int cedar_b200_pipe_open(const struct cedar_b200_config *cfg, const int *devices, int ndevices, int handles_per_device,
                         int gops_per_batch, cedar_b200_pipe **out)
{
    if (!cfg || !out || ndevices < 0 || handles_per_device < 0 || gops_per_batch < 0 || cfg->keyframe_interval <= 0)
        return -EINVAL;
    const int one = cfg->device;
    if (ndevices == 0 || !devices)
        devices = &one, ndevices = 1;
    if (handles_per_device == 0)
        handles_per_device = 2; // two chains per GPU fill the SMs one chain's wavefront kernels leave idle
    if (gops_per_batch == 0)
        gops_per_batch = 4;
    const auto t_open = std::chrono::steady_clock::now();
    cedar_b200_pipe *p = new cedar_b200_pipe();
    p->K = cfg->keyframe_interval;
    p->cap = gops_per_batch * p->K;
    p->w.resize((size_t)ndevices * handles_per_device);
    // worker order: device fastest, so that consecutive batches land on different GPUs.  Every worker opens its own handle
    // on its own thread: creating a CUDA context takes a second or two per device, and eight of them in a row is most of
    // the wall time of a short clip.
    std::vector<int> rc(p->w.size(), 0);
    {
        std::vector<std::thread> openers;
        for (size_t i = 0; i < p->w.size(); i++) {
            Worker &w = p->w[i];
            w.device = devices[i % (size_t)ndevices];
            openers.emplace_back([&, i] {
                Worker &wk = p->w[i];
                cedar_b200_config c = *cfg;
                c.device = wk.device;
                c.max_clip_frames = p->cap;
                cedar_b200_io io;
                rc[i] = cedar_b200_open(&c, &io, &wk.h);
                if (rc[i])
                    return;
                size_t fb = 0;
                wk.staging = (uint8_t *)cedar_b200_clip_input(wk.h, &fb);
                wk.sizes.resize((size_t)p->cap);
                wk.sse.resize((size_t)p->cap);
            });
        }
        for (auto &t : openers)
            t.join();
    }
    int r = 0;
    for (size_t i = 0; i < p->w.size(); i++)
        if (rc[i] && !r)
            r = rc[i];
    if (!r)
        cedar_b200_clip_input(p->w[0].h, &p->frame_bytes);
    if (r) {
        for (Worker &w : p->w)
            if (w.h)
                cedar_b200_close(w.h);
        delete p;
        return r;
    }
    for (size_t i = 0; i < p->w.size(); i++)
        p->w[i].thr = std::thread(worker_main, p, (int)i);
    if (getenv("CEDAR_B200_TRACE"))
        fprintf(stderr, "[trace] pipe_open: %zu workers x %d frames in %.1f ms\n", p->w.size(), p->cap,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_open).count());
    *out = p;
    return 0;
}

int cedar_b200_pipe_workers(cedar_b200_pipe *p) { return p ? (int)p->w.size() : 0; }

void *cedar_b200_pipe_acquire(cedar_b200_pipe *p, size_t *frame_bytes, int *capacity_frames)
{
    if (!p)
        return nullptr;
    std::unique_lock<std::mutex> lk(p->mu);
    if (p->acquired != p->submitted || p->finished || p->partial_seen)
        return nullptr; // one batch is filled at a time, in order; a short batch was the last one of the stream
    Worker &w = p->w[(size_t)(p->acquired % (long long)p->w.size())];
    p->cv.wait(lk, [&] { return w.state == FREE; });
    w.state = FILLING;
    p->acquired++;
    if (frame_bytes)
        *frame_bytes = p->frame_bytes;
    if (capacity_frames)
        *capacity_frames = p->cap;
    return w.staging;
}

int cedar_b200_pipe_submit(cedar_b200_pipe *p, int nframes)
{
    if (!p)
        return -EINVAL;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        if (p->acquired != p->submitted + 1 || nframes < 0 || nframes > p->cap)
            return -EINVAL;
        Worker &w = p->w[(size_t)(p->submitted % (long long)p->w.size())];
        if (nframes == 0) { // nothing read: give the slot back, the stream is over
            w.state = FREE;
            p->acquired--;
            p->finished = true;
        } else {
            if (p->partial_seen)
                return -EINVAL; // only the last batch may be short: batches start at GOP boundaries
            w.nframes = nframes;
            w.first_frame = (int)p->frames_submitted;
            p->frames_submitted += nframes;
            p->partial_seen = nframes < p->cap;
            w.state = QUEUED;
            p->submitted++;
        }
    }
    p->cv.notify_all();
    return 0;
}

int cedar_b200_pipe_finish(cedar_b200_pipe *p)
{
    if (!p)
        return -EINVAL;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->finished = true;
    }
    p->cv.notify_all();
    return 0;
}

long long cedar_b200_pipe_next(cedar_b200_pipe *p, const uint8_t **out, const int **frame_sizes, int *nframes,
                               const double **sse_y, int wait)
{
    if (!p)
        return -EINVAL;
    std::unique_lock<std::mutex> lk(p->mu);
    if (p->consuming >= 0) { // the batch handed out by the previous call is done with: its worker may go on
        p->w[(size_t)(p->consuming % (long long)p->w.size())].state = FREE;
        p->consuming = -1;
        p->cv.notify_all();
    }
    if (nframes)
        *nframes = 0;
    for (;;) {
        if (p->consumed < p->submitted) {
            Worker &w = p->w[(size_t)(p->consumed % (long long)p->w.size())];
            if (w.state == DONE) {
                w.state = CONSUMING;
                p->consuming = p->consumed++;
                if (out)
                    *out = w.out;
                if (frame_sizes)
                    *frame_sizes = w.sizes.data();
                if (sse_y)
                    *sse_y = w.sse.data();
                if (nframes)
                    *nframes = w.nframes;
                return w.total;
            }
        } else if (p->finished || !wait) {
            return 0; // drained (or nothing submitted yet)
        }
        if (!wait)
            return -EAGAIN;
        p->cv.wait(lk);
    }
}

int cedar_b200_pipe_release(cedar_b200_pipe *p)
{
    if (!p)
        return -EINVAL;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        if (p->consuming < 0)
            return 0;
        p->w[(size_t)(p->consuming % (long long)p->w.size())].state = FREE;
        p->consuming = -1;
    }
    p->cv.notify_all();
    return 0;
}

void cedar_b200_pipe_close(cedar_b200_pipe *p)
{
    if (!p)
        return;
    {
        std::unique_lock<std::mutex> lk(p->mu);
        // let running batches finish: a worker inside a clip call cannot be interrupted
        p->cv.wait(lk, [&] {
            for (Worker &w : p->w)
                if (w.state == QUEUED)
                    return false;
            return true;
        });
        p->stop = true;
    }
    p->cv.notify_all();
    for (Worker &w : p->w)
        if (w.thr.joinable())
            w.thr.join();
    for (Worker &w : p->w)
        if (w.h)
            cedar_b200_close(w.h);
    delete p;
}

} // extern "C"
