/*
 * oracle/ffcnn_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU restatement, "the oracle").
 *
 * A from-scratch plain-C restatement of the arithmetic on ffcnn's hot path, written
 * from the semantics of the reference (citations are file:line under /root/reference):
 *
 *   orc_groupconv   conv contract of conv.h:4-7; exact math = conv-v0.c:12-29
 *                   (accumulate in-channel -> kernel row -> kernel column, fp32, taps
 *                   outside the image skipped), epilogue act(sum*s + b) (conv-v0.c:27,
 *                   utils.h:15-23).  v6_quirk=1 additionally reproduces the two deviations
 *                   of the default build's depthwise 5x5 path (conv-v6.c:291-465):
 *                   output row oh-2 ignores kernel row 0 (conv-v6.c:422-441) and rows
 *                   0,1,oh-2,oh-1 accumulate column-outer / row-inner (conv-v6.c:323-333).
 *   orc_maxpool     clamped-window max, ffcnn.c:354-372,381-394
 *   orc_avgpool     clamped-window sum / fs^2, ffcnn.c:337-352
 *   orc_upsample    nearest, ffcnn.c:396-410
 *   orc_shortcut    act(a + b), ffcnn.c:418-423
 *   orc_net_input   BGR u8 -> planar RGB fp32, nearest resize to the top-left, ffcnn.c:259-289
 *   orc_fold_bn     BN fold into (scale, bias), ffcnn.c:222-233
 *   orc_yolo_decode candidate decode, ffcnn.c:438-474
 *   orc_nms         sort + greedy per-class min-area NMS + rescale, ffcnn.c:291-335
 *
 * Tensors are single-image planar CHW fp32 and filters are the packed rows
 * [ALIGN(fs*fs*ic/g,4) weights | scale, bias, mean, var] exactly as the reference
 * keeps them (ffcnn.c:150,218-234), so the same buffers can be handed to the compiled
 * reference (oracle/_ref) and to this file.
 *
 * Pinned against the reference itself: tests/test_oracle.py runs both on the same seeded
 * inputs (bit-exact against conv-v0/-O2; bit-exact against conv-v6/-O2 with v6_quirk=1)
 * and against the committed golden vectors in tests/golden/.
 * Build: -O2 -ffp-contract=off (no FMA contraction, no re-association).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_ALIGN4(x) (((x) + 3) & ~3)

typedef struct { int type; float score, x1, y1, x2, y2; } OrcBox;   /* == BBOX, ffcnn.h:29-32 */

static float orc_act(float v, int act)
{
    if (act == 1) return v > 0 ? v : 0;            /* relu   */
    if (act == 2) return v > 0 ? v : 0.1f * v;     /* leaky  */
    if (act == 3) return 1.0f / (1.0f + (float)exp(-v));
    return v;                                      /* linear / unknown(-1) */
}

/* is this the geometry that conv-v6.c:499 sends down its 5x5 depthwise fast path? */
static int orc_is_v6_dw5(int ic, int ig, int pad, int stride, int fs, int ow, int oh)
{
    return pad == 2 && fs == 5 && stride == 1 && ic / ig == 1 && oh >= 4 && ow >= 4;   /* below 4x4 the reference reads out of bounds (undefined) */
}

void orc_groupconv(const float *in, const float *flt, float *out,
                   int iw, int ih, int ic, int ig, int pad, int stride,
                   int fs, int fn, int ow, int oh, int oc, int act, int v6_quirk)
{
    const int cpg = ic / ig;                 /* input channels per group  */
    const int opg = oc / ig;                 /* output channels per group */
    const int row = ORC_ALIGN4(fs * fs * cpg) + 4;
    const int quirk = v6_quirk && orc_is_v6_dw5(ic, ig, pad, stride, fs, ow, oh);
    (void)fn;
    for (int o = 0; o < oc; o++) {
        const int    g  = o / opg;
        const float *w  = flt + (size_t)o * row;
        const float  s  = w[row - 4], b = w[row - 3];
        const float *gi = in + (size_t)g * cpg * iw * ih;
        float       *po = out + (size_t)o * ow * oh;
        for (int y = 0; y < oh; y++) {
            const int y0 = y * stride - pad;
            int j_lo = y0 < 0 ? -y0 : 0, j_hi = y0 + fs > ih ? ih - y0 : fs;
            const int edge_row = quirk && (y < 2 || y >= oh - 2);
            if (quirk && y == oh - 2) j_lo = 1;           /* conv-v6.c:426-437: kernel row 0 never read */
            for (int x = 0; x < ow; x++) {
                const int x0 = x * stride - pad;
                const int k_lo = x0 < 0 ? -x0 : 0, k_hi = x0 + fs > iw ? iw - x0 : fs;
                float sum = 0;
                if (!edge_row) {
                    for (int c = 0; c < cpg; c++)
                        for (int j = j_lo; j < j_hi; j++)
                            for (int k = k_lo; k < k_hi; k++)
                                sum += gi[((size_t)c * ih + (y0 + j)) * iw + x0 + k] * w[(c * fs + j) * fs + k];
                } else {                                   /* column-outer order of conv-v6.c:327-332 */
                    for (int k = k_lo; k < k_hi; k++)
                        for (int j = j_lo; j < j_hi; j++)
                            sum += gi[(size_t)(y0 + j) * iw + x0 + k] * w[j * fs + k];
                }
                float v = sum * s;
                v = v + b;
                po[y * ow + x] = orc_act(v, act);
            }
        }
    }
}

static void orc_window(int pos, int fs, int limit, int *lo, int *hi)
{
    int a = pos - (fs - 1) / 2, b = a + fs;
    *lo = a < 0 ? 0 : a; *hi = b > limit ? limit : b;
}

void orc_maxpool(const float *in, float *out, int w, int h, int c, int fs, int stride)
{
    const int ow = w / stride, oh = h / stride;   /* ffcnn.c:156-157; the loops below visit ceil(w/stride) columns */
    (void)ow; (void)oh;
    for (int ch = 0; ch < c; ch++) {
        const float *p = in + (size_t)ch * w * h;
        for (int iy = 0; iy < h; iy += stride)
            for (int ix = 0; ix < w; ix += stride) {
                int xa, xb, ya, yb; orc_window(ix, fs, w, &xa, &xb); orc_window(iy, fs, h, &ya, &yb);
                float m = p[ya * w + xa];
                for (int y = ya; y < yb; y++) for (int x = xa; x < xb; x++) if (m < p[y * w + x]) m = p[y * w + x];
                *out++ = m;
            }
    }
}

void orc_avgpool(const float *in, float *out, int w, int h, int c, int fs, int stride)
{
    for (int ch = 0; ch < c; ch++) {
        const float *p = in + (size_t)ch * w * h;
        for (int iy = 0; iy < h; iy += stride)
            for (int ix = 0; ix < w; ix += stride) {
                int xa, xb, ya, yb; orc_window(ix, fs, w, &xa, &xb); orc_window(iy, fs, h, &ya, &yb);
                float acc = 0;
                for (int y = ya; y < yb; y++) for (int x = xa; x < xb; x++) acc += p[y * w + x];
                *out++ = acc / (fs * fs);
            }
    }
}

void orc_upsample(const float *in, float *out, int w, int h, int c, int stride)
{
    const int ow = w * stride, oh = h * stride;
    for (int ch = 0; ch < c; ch++)
        for (int y = 0; y < oh; y++)
            for (int x = 0; x < ow; x++)
                out[((size_t)ch * oh + y) * ow + x] = in[((size_t)ch * h + y / stride) * w + x / stride];
}

void orc_shortcut(const float *a, const float *b, float *out, int n, int act)
{
    for (int i = 0; i < n; i++) out[i] = orc_act(a[i] + b[i], act);
}

void orc_net_input(const uint8_t *bgr, int w, int h, const float *mean, const float *norm,
                   float *out, int W, int H, int *s1_out, int *s2_out)
{
    int sw, sh, s1, s2;
    if (w * H > h * W) { sw = W; sh = W * h / w; s1 = w; s2 = sw; }
    else               { sh = H; sw = H * w / h; s1 = h; s2 = sh; }
    const int pitch = (w * 3 + 3) & ~3;
    float *r = out, *g = out + (size_t)W * H, *b = out + (size_t)2 * W * H;
    for (int i = 0; i < sh; i++)
        for (int j = 0; j < sw; j++) {
            const uint8_t *px = bgr + (size_t)(i * s1 / s2) * pitch + (j * s1 / s2) * 3;
            r[i * W + j] = (px[2] - mean[0]) * norm[0];
            g[i * W + j] = (px[1] - mean[1]) * norm[1];
            b[i * W + j] = (px[0] - mean[2]) * norm[2];
        }
    if (s1_out) *s1_out = s1;
    if (s2_out) *s2_out = s2;
}

/* scale[], bias[] in/out; mean[], var[] in. float add, double sqrt, float divide (ffcnn.c:230-231) */
void orc_fold_bn(float *scale, float *bias, const float *mean, const float *var, int n)
{
    for (int i = 0; i < n; i++) {
        scale[i] /= (float)sqrt(var[i] + 0.00001f);
        bias[i]  -= mean[i] * scale[i];
    }
}

/* head: CHW [3*(5+classes)][gh][gw]; anchors: 3 (w,h) pairs already mask-selected. Appends to boxes[n0..cap). */
int orc_yolo_decode(const float *head, int gw, int gh, int classes, const int *anchors,
                    float thresh, float scale_xy, int netw, int neth, OrcBox *boxes, int cap, int n0)
{
    const size_t plane = (size_t)gw * gh;
    int n = n0;
    for (int i = 0; i < gh; i++)
        for (int j = 0; j < gw; j++)
            for (int k = 0; k < 3; k++) {
                const float *cell = head + (size_t)k * (5 + classes) * plane + (size_t)i * gw + j;
                float bs = cell[4 * plane], cs = cell[5 * plane];
                int best = 0;
                for (int l = 1; l < classes; l++) { float v = cell[(5 + l) * plane]; if (cs < v) { cs = v; best = l; } }
                float conf = 1.0f / ((1.0f + (float)exp(-bs) * (1.0f + (float)exp(-cs))));
                if (!(conf >= thresh)) continue;
                float sx = 1.0f / (1.0f + (float)exp(-cell[0]));
                float sy = 1.0f / (1.0f + (float)exp(-cell[plane]));
                float cx = (j + sx) * netw / gw;
                float cy = (i + sy) * neth / gh;
                float bw = (float)exp(cell[2 * plane]) * anchors[2 * k]     * scale_xy;
                float bh = (float)exp(cell[3 * plane]) * anchors[2 * k + 1] * scale_xy;
                if (n < cap) {
                    boxes[n].type = best; boxes[n].score = conf;
                    boxes[n].x1 = cx - bw * 0.5f; boxes[n].y1 = cy - bh * 0.5f;
                    boxes[n].x2 = cx + bw * 0.5f; boxes[n].y2 = cy + bh * 0.5f;
                    n++;
                }
            }
    return n;
}

static int orc_by_score_desc(const void *a, const void *b)
{
    float sa = ((const OrcBox *)a)->score, sb = ((const OrcBox *)b)->score;
    return sa < sb ? 1 : sa > sb ? -1 : 0;
}

/* threshold 0.5 / min-area mode / rescale by s1/s2 are the arguments net_forward passes (ffcnn.c:519) */
int orc_nms(OrcBox *bx, int n, float threshold, int min_mode, int s1, int s2)
{
    if (!bx || n <= 0) return 0;
    qsort(bx, n, sizeof(OrcBox), orc_by_score_desc);
    /* "c" walks the survivors in score order; anything of the same class overlapping it is zeroed.
       The reference picks the next c as the first later box that was examined and not zeroed. */
    int c = 0;
    while (c >= 0 && c < n) {
        int next = -1;
        for (int j = c + 1; j < n; j++) {
            if (bx[j].score == 0) continue;
            int killed = 0;
            if (bx[c].type == bx[j].type) {
                float xa = bx[c].x1 > bx[j].x1 ? bx[c].x1 : bx[j].x1, ya = bx[c].y1 > bx[j].y1 ? bx[c].y1 : bx[j].y1;
                float xb = bx[c].x2 < bx[j].x2 ? bx[c].x2 : bx[j].x2, yb = bx[c].y2 < bx[j].y2 ? bx[c].y2 : bx[j].y2;
                float inter = (xa < xb && ya < yb) ? (xb - xa) * (yb - ya) : 0;
                float a1 = (bx[c].x2 - bx[c].x1) * (bx[c].y2 - bx[c].y1);
                float a2 = (bx[j].x2 - bx[j].x1) * (bx[j].y2 - bx[j].y1);
                float ratio = min_mode ? inter / (a1 < a2 ? a1 : a2) : inter / (a1 + a2 - inter);
                if (ratio > threshold) { bx[j].score = 0; killed = 1; }
            }
            if (!killed && next < 0) next = j;
        }
        c = next;
    }
    int m = 0;
    for (int i = 0; i < n; i++) {
        if (bx[i].score == 0) continue;
        OrcBox t = bx[i];
        t.x1 = t.x1 * s1 / s2; t.y1 = t.y1 * s1 / s2; t.x2 = t.x2 * s1 / s2; t.y2 = t.y2 * s1 / s2;
        bx[m++] = t;
    }
    memset(bx + m, 0, sizeof(OrcBox) * (n - m));
    return m;
}
