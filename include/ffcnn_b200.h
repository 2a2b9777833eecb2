/*
 * ffcnn_b200.h -- additive C-ABI of libffcnn_b200.so (plain pointers and sizes only).
 *
 * The reference has no batch dimension, no device and no multi-GPU notion; its whole public
 * surface is ffcnn.h:48-52 + conv.h:4-7 (kept, see include/ffcnn.h and include/conv.h).
 * The entry points below are what a reference-side binding adds to drive the same hot path
 * (net_input -> net_forward -> bbox_list, ffcnn.c:259-289,476-520) over BATCHES of frames on
 * a B200.  Each one names the reference code it stands in for.
 *
 * Conventions: int return = 0 on success, negative on failure with text in ffb_last_error().
 * All device work is enqueued on the net's stream (ffb_set_stream); calls that hand results
 * to the host synchronise that stream themselves.  Nothing here falls back to the CPU.
 */
#ifndef FFCNN_B200_EXT_H
#define FFCNN_B200_EXT_H

#include <stddef.h>
#include "ffcnn.h"

#ifdef __cplusplus
extern "C" {
#endif

#define FFB_E_OVERFLOW (-2)   /* ffb_detect_finish / ffb_collect: more yolo candidates than the device buffer holds (never a silent truncation) */

const char *ffb_last_error(void);
int         ffb_device_count(void);                      /* 0 without a usable CUDA device */

/* ---- load, split in two so the host half is testable without a GPU ---------------------- */

/* Host half of net_load (ffcnn.c:114-239): cfg parse, weights read, filter packing, BN fold.
 * weightsfile may be NULL / missing -> zero weights, as the reference (ffcnn.c:213-220). */
NET *ffb_net_parse(const char *cfgfile, const char *weightsfile, int inputw, int inputh);

/* Device half: pick `device`, upload NET.weight_buf, derive per-kernel weight layouts, plan the
 * activation arena for up to max_batch frames.  net_load() = ffb_net_parse + ffb_net_attach(net,
 * $FFCNN_DEVICE or 0, 1).  Re-attaching with a larger max_batch re-plans. */
int  ffb_net_attach(NET *net, int device, int max_batch);

/* Device copy of NET.weight_buf (packed rows, ffcnn.c:218-234), weight_size floats.  Exposed so
 * a multi-GPU frontend can broadcast rank 0's weights into it (NCCL) and then call
 * ffb_commit_weights() to rebuild the kernel-side layouts from it. */
void *ffb_packed_weights_device(NET *net, size_t *nfloats);
int   ffb_commit_weights(NET *net);

/* Options (name, value): "dw5_exact" 0 = reproduce conv-v6's dropped kernel row on output row
 * oh-2 of the 5x5 depthwise path (conv-v6.c:422-441; default, the named oracle), 1 = exact math
 * (conv-v0); "pw_mode" 0 = auto, 1 = fp32 FFMA everywhere, 2 = tcgen05 3xTF32 where eligible,
 * 3 = tcgen05 1xTF32; "dw_mode" 0 = TMA-fed shared-memory stencil for the depthwise layers (default), 1 = the
 * register-window kernel fed by plain loads; "graph" 1 = replay a captured CUDA graph (default), 0 = plain launches;
 * "keep_all" 1 = layer-by-layer plan, every layer output gets its own buffer (ffb_layer_output works for every layer),
 * 2 = the fused plan with one private buffer per surviving tensor; "fuse_block" 1 (default) = run the expand -> depthwise ->
 * project [-> shortcut] chains where it pays as ONE kernel each, 2 = every supported chain, 0 = off;
 * "fuse_tail" 1 (default) = SPP pools + route as one kernel, upsample writes into its concat tensor; "fuse_input" 1 (default) = when
 * the frames already have the net's size, the stem kernel reads the u8 frames itself (net_input fused, no fp32 input
 * tensor) -- the frame buffer handed to ffb_input_u8 must then stay valid until ffb_forward's work has completed. */
int  ffb_set_option(NET *net, const char *name, int value);
int  ffb_get_option(NET *net, const char *name);

int   ffb_set_stream(NET *net, void *cuda_stream);       /* cudaStream_t; NULL = engine's own */
void *ffb_get_stream(NET *net);
int   ffb_sync(NET *net);

/* ---- batched hot path -------------------------------------------------------------------- */

/* Batched net_input (ffcnn.c:259-289): n frames of w x h BGR u8, rows top-down, `pitch` bytes
 * per row, frame stride h*pitch.  Nearest-neighbour fit to the top-left of the net input,
 * (px - mean[c]) * norm[c], rest zero.  frames_on_device != 0: `frames` is a device pointer
 * (resident input); otherwise it is host memory (pinned for async copies) and the H2D copy is
 * enqueued here. */
int  ffb_input_u8(NET *net, const unsigned char *frames, int n, int w, int h, int pitch,
                  const float *mean, const float *norm, int frames_on_device);

/* n host CHW fp32 tensors [n][c][H][W] as the network input (what layer_list[0].data holds). */
int  ffb_input_chw(NET *net, const float *chw, int n, int s1, int s2);

/* The layer loop of net_forward (ffcnn.c:488-518) for the current batch, asynchronous. */
int  ffb_forward(NET *net);

/* yolo decode + NMS (ffcnn.c:438-474,298-335) for every frame of the batch: the GPU filters
 * candidates, the host finishes with the reference's exact libm arithmetic.  Synchronises. */
int  ffb_detect(NET *net);
/* The two halves of ffb_detect, for callers that overlap host work with the GPU: enqueue = filter kernels
 * + async copy of the candidate count (returns the number of yolo heads); finish = wait, fetch, decode, NMS. */
int  ffb_detect_enqueue(NET *net);
int  ffb_detect_finish(NET *net);
long ffb_last_d2h_bytes(NET *net);                        /* bytes the last ffb_detect_finish copied to the host */

/* Boxes of frame `frame` after ffb_detect; pointer valid until the next ffb_detect. */
int  ffb_boxes(NET *net, int frame, BBOX **boxes);
int  ffb_raw_boxes(NET *net, int frame, BBOX **boxes);   /* pre-NMS candidates in reference scan order */

/* ffb_input_u8 + ffb_forward + ffb_detect: the end-to-end call (host frames in, boxes out). */
int  ffb_detect_batch_u8(NET *net, const unsigned char *frames_host, int n, int w, int h, int pitch,
                         const float *mean, const float *norm);

/* Pipelined form of the same call for streams of batches: ffb_submit_u8 starts the H2D copy of a batch on a copy
 * stream (at most three batches in flight); ffb_collect runs forward + detect for the oldest submitted batch and leaves its
 * boxes readable with ffb_boxes.  submit(b0); loop { submit(b_next); collect(); read boxes; } overlaps the PCIe copy of
 * the next batch with the GPU work and the host decode of the current one.  frames_host must stay valid (and should be
 * pinned) until the matching ffb_collect returns. */
int  ffb_submit_u8(NET *net, const unsigned char *frames_host, int n, int w, int h, int pitch,
                   const float *mean, const float *norm);
int  ffb_collect(NET *net);

/* ---- multi-GPU frontend (one box) --------------------------------------------------------- */

/* The reference runs one image on one thread (ffcnn.c:476-520) and has no device notion; frames are independent, so the
 * batched frontend is data parallel: one NET per device, each with its own host thread, stream and CUDA graph; a batch of
 * n frames is cut into contiguous shards (device g of G gets frames [g*n/G, (g+1)*n/G)); nothing crosses devices per
 * forward pass.  Load: the first device's NET reads the weights file (ffcnn.c:211-239), every other device receives the
 * packed buffer NET.weight_buf (ffcnn.c:150: weight_size floats) through ONE ncclBroadcast over NVLink/NVSwitch
 * (libnccl.so.2 is dlopen'ed -- $FFCNN_NCCL_LIB overrides -- so a single-device program never needs it) and rebuilds its
 * kernel-side layouts.  devices == NULL / ndev <= 0: every visible device.  A device may be listed twice (two replicas
 * sharing one GPU; the weights then travel by a device-to-device copy).  Calls are made from ONE caller thread. */
typedef struct ffb_multi ffb_multi;
ffb_multi *ffb_multi_create(const char *cfgfile, const char *weightsfile, int inputw, int inputh,
                            const int *devices, int ndev, int max_batch_per_device);
void  ffb_multi_destroy(ffb_multi *m);
int   ffb_multi_devices(ffb_multi *m);
NET  *ffb_multi_net(ffb_multi *m, int index);            /* the per-device NET (options, inspection) */
long  ffb_multi_broadcast_bytes(ffb_multi *m);           /* bytes the NCCL weight broadcast moved (0 for one device) */
/* Sharded ffb_detect_batch_u8 / ffb_submit_u8 / ffb_collect: same arguments and pipelining rules, n = frames in total. */
int   ffb_multi_detect_u8(ffb_multi *m, const unsigned char *frames_host, int n, int w, int h, int pitch,
                          const float *mean, const float *norm);
int   ffb_multi_submit_u8(ffb_multi *m, const unsigned char *frames_host, int n, int w, int h, int pitch,
                          const float *mean, const float *norm);
int   ffb_multi_collect(ffb_multi *m);
int   ffb_multi_boxes(ffb_multi *m, int frame, BBOX **boxes);   /* frame = index in the whole batch */

/* ---- inspection / measurement ------------------------------------------------------------ */

/* Copy the output of layer `layer` for frame `frame` to host as CHW fp32 (the reference layout).
 * Needs option keep_all=1 (or 2: then only tensors the fused plan materialises exist, others return 0) set before
 * ffb_forward.  Returns the number of floats. */
long ffb_layer_output(NET *net, int layer, int frame, float *chw_host, long capacity);

/* Time every layer of the current batch with CUDA events: ms[i] = mean over `reps` launches of
 * layer i's kernel(s) (0 for aliases: dropout, single-input route).  flush_l2 != 0 evicts L2
 * between launches. */
int  ffb_layer_times(NET *net, float *ms, int nlayers, int reps, int flush_l2);

/* Algorithmic bytes / flops per frame of layer i as defined in SURVEY 8: fp32 input + output
 * activations + that layer's weights, each touched once. */
int  ffb_layer_cost(NET *net, int layer, double *bytes, double *flops, char *kernel_name, int name_cap);

int  ffb_launches_per_forward(NET *net);                 /* kernels enqueued by one ffb_forward */

/* Measured dense tcgen05.mma.kind::tf32 rate of the current device in TFLOP/s (every SM issuing M=128 N=256 K=8 MMAs
 * back to back from shared memory): the tensor-pipe denominator of the roofline report.  ~10 ms. */
int  ffb_measure_tf32_peak(double *tflops);

/* ---- single operator on device tensors (configs 3 and 4 of BASELINE.json; op-level parity) - */

typedef struct ffb_conv ffb_conv;
/* Same contract as groupconv (conv.h:4-7) but batched NHWC on device, weights uploaded once.
 * packed_filter: host, fn rows of ALIGN(fs*fs*ic/groups,4)+4 floats.  flags: bit0 = dw5_exact,
 * bits 8..15 = pw_mode, bits 16..23 = dw_mode. */
ffb_conv *ffb_conv_create(const float *packed_filter, int ic, int groups, int pad, int stride,
                          int fs, int fn, int activation, int flags);
void      ffb_conv_destroy(ffb_conv *op);
int       ffb_conv_run(ffb_conv *op, const float *in_nhwc_dev, float *out_nhwc_dev,
                       int n, int ih, int iw, void *cuda_stream);
const char *ffb_conv_kernel_name(ffb_conv *op);

void *ffb_dev_alloc(size_t bytes);
void  ffb_dev_free(void *p);
int   ffb_copy_h2d(void *dst_dev, const void *src_host, size_t bytes);
int   ffb_copy_d2h(void *dst_host, const void *src_dev, size_t bytes);
void *ffb_host_alloc_pinned(size_t bytes);
void *ffb_host_alloc_pinned_wc(size_t bytes);            /* write-combined pinned memory (host writes only; free with ffb_host_free_pinned) */
void  ffb_host_free_pinned(void *p);
/* layout helpers on device: [n][c][h][w] <-> [n][h][w][c] */
int   ffb_chw_to_nhwc(const float *src_dev, float *dst_dev, int n, int c, int h, int w, void *cuda_stream);
int   ffb_nhwc_to_chw(const float *src_dev, float *dst_dev, int n, int c, int h, int w, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
