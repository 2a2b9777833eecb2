/*
 * conv.h -- the reference's operator seam, exported by libffcnn_b200.so.
 *
 * Replaces /root/reference/conv.h:4-7.  The reference selects one of conv-v0.c..conv-v6.c at
 * link time (build.sh:48); linking libffcnn_b200.so instead of a conv-vN.c gives the
 * unmodified reference ffcnn.c a B200-backed convolution (host CHW in, host CHW out).
 *
 *   datai  CHW fp32 input  [ic][ih][iw]            (host)
 *   dataf  packed filters: fn rows of ALIGN(fs*fs*ic/ig, 4) + 4 floats,
 *          row = [weights..., zero pad, scale, bias, mean, var]   (ffcnn.c:218-234)
 *   datao  CHW fp32 output [oc][oh][ow], caller-allocated (ffcnn.c:490)
 *   activation  0 linear, 1 relu, 2 leaky(0.1), anything else linear (utils.h:8-23)
 *   gc_buffer / gc_bufsize  the caller's scratch slot (ffcnn.h:43-44); the GPU path never
 *          touches it, so it stays NULL / 0 and net_free's free() remains valid.
 * Errors: returns void like the reference; on a CUDA failure it prints to stderr and leaves
 * datao unwritten (conv-v6.c:509 behaves the same way on malloc failure).
 */
#ifndef FFCNN_B200_CONV_H
#define FFCNN_B200_CONV_H
#ifdef __cplusplus
extern "C" {
#endif

void groupconv(float *datai, float *dataf, float *datao,
               int iw, int ih, int ic, int ig, int ipad, int istride,
               int fs, int fn, int ow, int oh, int oc, int activation,
               float **gc_buffer, int *gc_bufsize);

#ifdef __cplusplus
}
#endif
#endif
