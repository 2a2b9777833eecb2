/*
 * ffb_internal.h -- private glue between the host C side (cfg/weights/io/decode) and the
 * CUDA engine.  Not installed; the public surface is the headers under include/.
 */
#ifndef FFB_INTERNAL_H
#define FFB_INTERNAL_H

#include <stddef.h>
#include "../../include/ffcnn.h"
#include "../../include/ffcnn_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define FFB_ALIGN(x, n) (((x) + ((n) - 1)) & ~((n) - 1))
#define FFB_MAGIC 0x46464232u          /* "FFB2" */

enum { FFB_ACT_LINEAR = 0, FFB_ACT_RELU = 1, FFB_ACT_LEAKY = 2 };

struct ffb_engine;                      /* CUDA side, engine.cu */
#define FFB_SLOTS 3                     /* batches ffb_submit_u8 accepts before ffb_collect must be called (frame slots, detection sets) */

/* The object net_load returns: the ABI-visible NET first, private state after it. */
typedef struct ffb_net {
    NET                pub;
    unsigned           magic;
    struct ffb_engine *engine;          /* NULL until ffb_net_attach */
    int                input_w, input_h, input_c;
    float              in_mean[3], in_norm[3];
} ffb_net;

static inline ffb_net *ffb_from_pub(NET *n) { return (ffb_net *)n; }

void ffb_set_error(const char *fmt, ...);

/* one candidate as the GPU filter emits it: everything the exact host decode needs */
typedef struct {
    int   frame;
    int   key;          /* ((head_index * cells + cell) * 3 + anchor): restores the reference's scan order */
    int   cls;          /* first arg-max over the class logits (ffcnn.c:446-450) */
    float bs, cs;       /* objectness logit, best class logit */
    float tx, ty, tw, th;
} ffb_candidate;

/* host_decode.c */
int  ffb_decode_candidate(const LAYER *yolo, int netw, int neth, int gw, int gh,
                          int cell, int anchor, const ffb_candidate *c, BBOX *out);
int  ffb_decode_head_chw(const LAYER *yolo, const float *head_chw, int gw, int gh, int netw, int neth,
                         BBOX *boxes, int n0, int cap);
int  ffb_nms(BBOX *boxes, int n, float threshold, int min_mode, int s1, int s2);
void ffb_fit_geometry(int w, int h, int W, int H, int *sw, int *sh, int *s1, int *s2);

/* engine.cu: per-device launch state (cudaFuncSetAttribute is per (function, device); several nets on different devices
 * may share one process).  A ffb_smem_cfg is a zero-initialised static next to the kernel it describes. */
#define FFB_MAX_DEVICES 64
typedef struct { size_t bytes[FFB_MAX_DEVICES]; } ffb_smem_cfg;
int  ffb_num_sms(void);                                                  /* SM count of the current device */
int  ffb_ensure_smem(const void *func, size_t smem, ffb_smem_cfg *cfg);  /* raise func's dynamic-smem limit on the current device if needed */

void ffb_engine_destroy(struct ffb_engine *e);
int  ffb_engine_forward_single(ffb_net *net);     /* net_forward(): layer_list[0].data -> bbox_list */

#ifdef __cplusplus
}
#endif
#endif
