/*
 * multi.cu -- the batched-image frontend over the GPUs of one box, in C (include/ffcnn_b200.h "multi-GPU").
 *
 * The reference has no device or multi-GPU notion: one NET, one image, one thread (ffcnn.c:476-520).  Frames are
 * independent, so the frontend is data parallel with nothing on the wire per forward pass:
 *
 *   ffb_multi_create   one NET per device.  The first one reads the weights file (ffcnn.c:211-239); the others are parsed
 *                      without weights and receive the packed buffer (NET.weight_buf, ffcnn.c:150) by ONE ncclBroadcast
 *                      from device 0 straight into their device copies (NVLink / NVSwitch), then rebuild their kernel-side
 *                      layouts (ffb_commit_weights).  NCCL is loaded with dlopen("libnccl.so.2"): no link-time dependency,
 *                      and a single-device frontend never touches it.  Two entries naming the SAME device (NCCL refuses
 *                      duplicates) get a device-to-device copy instead.
 *   one host thread per device, each driving its own NET / stream / CUDA graph through the single-device entry points
 *   ffb_multi_detect_u8 / submit_u8 / collect   contiguous frame shards: device g of G gets frames [g*n/G, (g+1)*n/G)
 *   ffb_multi_boxes    frame-ordered view over the per-device results
 *
 * Every call fans the same single-device call out to the workers and returns the first error.
 */
#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

#include "ffb_internal.h"

namespace {

/* ---- the five NCCL entry points the frontend needs, resolved at run time ---- */
typedef struct ncclComm *ncclComm_t;
struct Nccl {
    void *h = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
    bool load()
    {
        if (h) return true;
        const char *names[] = { getenv("FFCNN_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
        for (const char *n : names) { if (n && (h = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break; }
        if (!h) { ffb_set_error("cannot load libnccl.so.2 (set FFCNN_NCCL_LIB): %s", dlerror()); return false; }
#define SYM(f) *(void **)(&f) = dlsym(h, "nccl" #f); if (!f) { ffb_set_error("libnccl lacks nccl" #f); return false; }
        SYM(CommInitAll) SYM(CommDestroy) SYM(Broadcast) SYM(GroupStart) SYM(GroupEnd) SYM(GetErrorString) SYM(GetVersion)
#undef SYM
        return true;
    }
};
Nccl g_nccl;
const int NCCL_FLOAT = 7;               /* ncclFloat32 (nccl.h: ncclDataType_t) */

enum Op { OP_NONE, OP_DETECT, OP_SUBMIT, OP_COLLECT, OP_QUIT };

struct Worker {
    NET *net = nullptr; int device = 0, index = 0;
    std::thread th; std::mutex mu; std::condition_variable cv;
    Op op = OP_NONE; bool done = true; int rc = 0; std::string err;
    const unsigned char *frames = nullptr; int n = 0, w = 0, h = 0, pitch = 0; const float *mean = nullptr, *norm = nullptr;
    int last_n = 0;                         /* frames of the batch whose boxes are currently readable */
    std::vector<int> inflight;              /* frame counts of submitted, not yet collected batches */
};

} // namespace

struct ffb_multi {
    std::vector<Worker *> workers;
    std::vector<int> first;                 /* first[g] = global index of worker g's first frame in the readable batch */
    std::vector<int> inflight_total;
    int total = 0;
    size_t bcast_bytes = 0; int nccl_version = 0;
};

static void worker_loop(Worker *w)
{
    cudaSetDevice(w->device);
    for (;;) {
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return !w->done; });
        const Op op = w->op;
        lk.unlock();
        int rc = 0;
        if (op == OP_QUIT) { lk.lock(); w->done = true; w->cv.notify_all(); return; }
        if (w->n > 0 || op == OP_COLLECT) {
            if (op == OP_DETECT)       { rc = ffb_detect_batch_u8(w->net, w->frames, w->n, w->w, w->h, w->pitch, w->mean, w->norm); if (rc == 0) w->last_n = w->n; }
            else if (op == OP_SUBMIT)  { rc = ffb_submit_u8(w->net, w->frames, w->n, w->w, w->h, w->pitch, w->mean, w->norm); if (rc == 0) w->inflight.push_back(w->n); }
            else if (op == OP_COLLECT) {
                if (w->inflight.empty()) w->last_n = 0;                 /* this device had no frames in that batch */
                else if (w->inflight.front() == 0) { w->last_n = 0; w->inflight.erase(w->inflight.begin()); }
                else { rc = ffb_collect(w->net); if (rc == 0) w->last_n = w->inflight.front(); w->inflight.erase(w->inflight.begin()); }
            }
        } else if (op == OP_DETECT) w->last_n = 0;
        else if (op == OP_SUBMIT) w->inflight.push_back(0);
        lk.lock();
        w->rc = rc; w->err = rc ? ffb_last_error() : "";
        w->done = true;
        w->cv.notify_all();
    }
}

static void post(Worker *w, Op op) { std::lock_guard<std::mutex> lk(w->mu); w->op = op; w->done = false; w->cv.notify_all(); }
static int  wait_done(Worker *w) { std::unique_lock<std::mutex> lk(w->mu); w->cv.wait(lk, [&] { return w->done; }); return w->rc; }

static int run_all(ffb_multi *m, Op op)
{
    for (Worker *w : m->workers) post(w, op);
    int rc = 0;
    for (Worker *w : m->workers) { const int r = wait_done(w); if (r && !rc) { rc = r; ffb_set_error("device %d: %s", w->device, w->err.c_str()); } }
    return rc;
}

extern "C" void ffb_multi_destroy(ffb_multi *m)
{
    if (!m) return;
    for (Worker *w : m->workers) {
        if (w->th.joinable()) { post(w, OP_QUIT); w->th.join(); }
        if (w->net) net_free(w->net);
        delete w;
    }
    delete m;
}

extern "C" ffb_multi *ffb_multi_create(const char *cfgfile, const char *weightsfile, int inputw, int inputh,
                                       const int *devices, int ndev, int max_batch_per_device)
{
    const int have = ffb_device_count();
    if (have <= 0) { ffb_set_error("no CUDA device available: libffcnn_b200 has no CPU fallback"); return nullptr; }
    std::vector<int> devs;
    if (!devices || ndev <= 0) for (int d = 0; d < have; d++) devs.push_back(d);
    else devs.assign(devices, devices + ndev);
    for (int d : devs) if (d < 0 || d >= have) { ffb_set_error("ffb_multi_create: device %d out of range (%d devices)", d, have); return nullptr; }
    ffb_multi *m = new ffb_multi();
    const int G = (int)devs.size();
    for (int g = 0; g < G; g++) {
        Worker *w = new Worker(); w->device = devs[g]; w->index = g;
        m->workers.push_back(w);
        /* only the first NET reads the weights file; the others get the packed buffer over NCCL below */
        w->net = ffb_net_parse(cfgfile, g == 0 ? weightsfile : nullptr, inputw, inputh);
        if (!w->net || ffb_net_attach(w->net, devs[g], std::max(1, max_batch_per_device)) != 0) { ffb_multi_destroy(m); return nullptr; }
    }
    if (G > 1) {
        size_t nfl = 0, nf2 = 0;
        void *src = ffb_packed_weights_device(m->workers[0]->net, &nfl);
        std::vector<int> distinct;                                         /* NCCL wants each device once */
        std::vector<int> owner(G, -1);
        for (int g = 0; g < G; g++) {
            int at = -1;
            for (size_t k = 0; k < distinct.size(); k++) if (devs[distinct[k]] == devs[g]) at = (int)k;
            if (at < 0) { distinct.push_back(g); owner[g] = g; } else owner[g] = distinct[at];
        }
        bool ok = src != nullptr;
        if (ok && distinct.size() > 1) {
            ok = g_nccl.load();
            std::vector<ncclComm_t> comms(distinct.size(), nullptr);
            std::vector<int> dl; for (int g : distinct) dl.push_back(devs[g]);
            int r = 0;
            if (ok && (r = g_nccl.CommInitAll(comms.data(), (int)dl.size(), dl.data())) != 0) { ffb_set_error("ncclCommInitAll: %s", g_nccl.GetErrorString(r)); ok = false; }
            if (ok) {
                g_nccl.GetVersion(&m->nccl_version);
                g_nccl.GroupStart();
                for (size_t k = 0; k < distinct.size() && ok; k++) {
                    NET *net = m->workers[distinct[k]]->net;
                    void *dst = ffb_packed_weights_device(net, &nf2);
                    cudaSetDevice(dl[k]);
                    if (!dst || nf2 != nfl) { ffb_set_error("ffb_multi_create: weight buffers differ between devices"); ok = false; break; }
                    r = g_nccl.Broadcast(k == 0 ? src : dst, dst, nfl, NCCL_FLOAT, 0, comms[k], (cudaStream_t)ffb_get_stream(net));
                    if (r != 0) { ffb_set_error("ncclBroadcast: %s", g_nccl.GetErrorString(r)); ok = false; }
                }
                r = g_nccl.GroupEnd();
                if (ok && r != 0) { ffb_set_error("ncclGroupEnd: %s", g_nccl.GetErrorString(r)); ok = false; }
                for (size_t k = 0; k < distinct.size(); k++) ffb_sync(m->workers[distinct[k]]->net);
                m->bcast_bytes = nfl * sizeof(float);
            }
            for (ncclComm_t c : comms) if (c) g_nccl.CommDestroy(c);
        }
        for (int g = 1; g < G && ok; g++) {
            Worker *w = m->workers[g];
            if (owner[g] != g) {                                           /* same device as an earlier entry: device-to-device copy */
                void *from = ffb_packed_weights_device(m->workers[owner[g]]->net, &nf2), *to = ffb_packed_weights_device(w->net, &nf2);
                cudaSetDevice(w->device);
                ok = from && to && cudaMemcpy(to, from, nfl * sizeof(float), cudaMemcpyDeviceToDevice) == cudaSuccess;
                if (!ok) ffb_set_error("ffb_multi_create: device-to-device weight copy failed");
            }
            if (ok && ffb_commit_weights(w->net) != 0) ok = false;
        }
        if (!ok) { ffb_multi_destroy(m); return nullptr; }
    }
    for (Worker *w : m->workers) w->th = std::thread(worker_loop, w);
    m->first.assign(G, 0);
    return m;
}

extern "C" int ffb_multi_devices(ffb_multi *m) { return m ? (int)m->workers.size() : 0; }
extern "C" NET *ffb_multi_net(ffb_multi *m, int g) { return m && g >= 0 && g < (int)m->workers.size() ? m->workers[g]->net : nullptr; }
extern "C" long ffb_multi_broadcast_bytes(ffb_multi *m) { return m ? (long)m->bcast_bytes : 0; }

/* contiguous shards: device g of G gets frames [g*n/G, (g+1)*n/G) */
static void shard(ffb_multi *m, const unsigned char *frames, int n, int w, int h, int pitch, const float *mean, const float *norm)
{
    const int G = (int)m->workers.size();
    for (int g = 0; g < G; g++) {
        Worker *wk = m->workers[g];
        const long lo = (long)g * n / G, hi = (long)(g + 1) * n / G;
        wk->frames = frames + (size_t)lo * h * pitch; wk->n = (int)(hi - lo); wk->w = w; wk->h = h; wk->pitch = pitch; wk->mean = mean; wk->norm = norm;
    }
}

static void publish(ffb_multi *m, int n)
{
    int at = 0;
    for (size_t g = 0; g < m->workers.size(); g++) { m->first[g] = at; at += m->workers[g]->last_n; }
    m->total = n;
}

extern "C" int ffb_multi_detect_u8(ffb_multi *m, const unsigned char *frames_host, int n, int w, int h, int pitch, const float *mean, const float *norm)
{
    if (!m || !frames_host || n < 1) { ffb_set_error("ffb_multi_detect_u8: bad arguments"); return -1; }
    shard(m, frames_host, n, w, h, pitch, mean, norm);
    const int rc = run_all(m, OP_DETECT);
    if (rc == 0) publish(m, n);
    return rc;
}

extern "C" int ffb_multi_submit_u8(ffb_multi *m, const unsigned char *frames_host, int n, int w, int h, int pitch, const float *mean, const float *norm)
{
    if (!m || !frames_host || n < 1) { ffb_set_error("ffb_multi_submit_u8: bad arguments"); return -1; }
    if (m->inflight_total.size() >= FFB_SLOTS) { ffb_set_error("ffb_multi_submit_u8: %d batches already in flight, call ffb_multi_collect first", FFB_SLOTS); return -1; }
    shard(m, frames_host, n, w, h, pitch, mean, norm);
    const int rc = run_all(m, OP_SUBMIT);
    if (rc == 0) m->inflight_total.push_back(n);
    return rc;
}

extern "C" int ffb_multi_collect(ffb_multi *m)
{
    if (!m || m->inflight_total.empty()) { ffb_set_error("ffb_multi_collect: nothing submitted"); return -1; }
    const int rc = run_all(m, OP_COLLECT);
    const int n = m->inflight_total.front(); m->inflight_total.erase(m->inflight_total.begin());
    if (rc == 0) publish(m, n);
    return rc;
}

extern "C" int ffb_multi_boxes(ffb_multi *m, int frame, BBOX **boxes)
{
    if (!m || frame < 0 || frame >= m->total) { ffb_set_error("ffb_multi_boxes: frame %d out of range", frame); return -1; }
    for (size_t g = 0; g < m->workers.size(); g++) {
        Worker *w = m->workers[g];
        if (frame >= m->first[g] && frame < m->first[g] + w->last_n) return ffb_boxes(w->net, frame - m->first[g], boxes);
    }
    ffb_set_error("ffb_multi_boxes: frame %d not found", frame);
    return -1;
}
