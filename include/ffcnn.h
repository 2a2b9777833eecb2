/*
 * ffcnn.h -- the reference's public network API, re-declared for the B200 build.
 *
 * Drop-in boundary: this header replaces /root/reference/ffcnn.h.  A caller
 * compiled against the reference header links against libffcnn_b200.so unchanged:
 *   - the five entry points have the reference's names and signatures (ffcnn.h:48-52),
 *     plus net_profile which the reference exports without declaring (ffcnn.c:550);
 *   - LAYER / BBOX / NET have the reference's field order and types (ffcnn.h:16-46),
 *     because callers read results straight out of NET (ffcnn.c:583-586).
 *
 * What differs behind the boundary: tensors live in B200 HBM as batched NHWC fp32 and
 * every layer runs as a hand-written sm_100a kernel; LAYER.data is only populated for
 * layer 0 (the host-side input tensor net_input fills) and bbox_list is its own buffer
 * instead of aliasing that tensor (ffcnn.c:243-244) -- see DESIGN.md.
 * There is no CPU fallback: net_load returns NULL (and says why on stderr) without a GPU.
 */
#ifndef FFCNN_B200_FFCNN_H
#define FFCNN_B200_FFCNN_H

#ifdef __cplusplus
extern "C" {
#endif

/* layer kinds, numbering of ffcnn.h:4-14 (TOTOAL is the reference's spelling) */
enum {
    LAYER_TYPE_CONV     = 0,
    LAYER_TYPE_AVGPOOL  = 1,
    LAYER_TYPE_MAXPOOL  = 2,
    LAYER_TYPE_UPSAMPLE = 3,
    LAYER_TYPE_DROPOUT  = 4,
    LAYER_TYPE_SHORTCUT = 5,
    LAYER_TYPE_ROUTE    = 6,
    LAYER_TYPE_YOLO     = 7,
    LAYER_TYPE_TOTOAL   = 8
};

/* entry i describes layer i and the geometry of its INPUT; entry i+1 its output (ffcnn.c:123-130) */
typedef struct {
    int    type, refcnt;
    float *data;                 /* host tensor: only layer_list[0].data (CHW fp32 input) is used here */
    float *filter;               /* this layer's rows inside NET.weight_buf (packed, ffcnn.c:218-234) */
    int    w, h, c, pad, stride, fn, fs, groups;
    int    batchnorm, activation;
    int    depend_list[4];
    int    depend_num;

    int    class_num;
    int    anchor_list[3][2];
    float  ignore_thres, scale_x_y;
} LAYER;

typedef struct {
    int   type;                  /* class index */
    float score, x1, y1, x2, y2; /* source-image pixels after net_forward */
} BBOX;

typedef struct {
    LAYER *layer_list;
    int    layer_num;
    BBOX  *bbox_list;
    int    bbox_num;
    int    bbox_max;
    int    s1, s2;               /* box rescale ratio source/net set by net_input */
    int    weight_size;          /* floats in weight_buf */
    float *weight_buf;           /* packed filters: [ALIGN(k*k*c/g,4) w | scale bias mean var] per filter */
    float *cnntempbuf;           /* unused by the GPU path; kept NULL so net_free's free() is safe */
    int    cnnbufsize;
    int    timeused[LAYER_TYPE_TOTOAL];
} NET;

NET *net_load   (char *cfgfile, char *weightsfile, int inputw, int inputh);
void net_free   (NET *net);
void net_input  (NET *net, unsigned char *bgr, int w, int h, float *mean, float *norm);
void net_forward(NET *net);
void net_dump   (NET *net);
void net_profile(NET *net);

#ifdef __cplusplus
}
#endif
#endif
