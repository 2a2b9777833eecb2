/*
 * host_decode.c -- "detection boxes out": yolo candidate decode and NMS on the host.
 *
 * Stays in host C on purpose ("yolo decode + NMS"): the reference's numbers come out
 * of double-precision libm exp() applied to float logits and rounded back to float
 * (ffcnn.c:451,457-460; utils.h:21).  Re-running exactly that arithmetic here, on the handful of
 * candidates the GPU filter lets through, keeps scores and boxes bit-identical to what the
 * reference would compute from the same head tensor.
 *
 *   ffb_decode_candidate / ffb_decode_head_chw ... ffcnn.c:438-474
 *   ffb_nms ...................................... ffcnn.c:291-335 (called as nms(.., 0.5f, 1, s1, s2), 519)
 *   ffb_fit_geometry ............................. the resize arithmetic of net_input, ffcnn.c:267-273
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "ffb_internal.h"

static float logistic(float v) { return 1.0f / (1.0f + (float)exp(-v)); }

static float confidence_of(float bs, float cs)
{
    /* not sigmoid(bs)*sigmoid(cs): the reference's own formula, reproduced as written (ffcnn.c:451) */
    return 1.0f / ((1.0f + (float)exp(-bs) * (1.0f + (float)exp(-cs))));
}

int ffb_decode_candidate(const LAYER *yolo, int netw, int neth, int gw, int gh,
                         int cell, int anchor, const ffb_candidate *c, BBOX *out)
{
    const int row = cell / gw, col = cell % gw;
    float conf = confidence_of(c->bs, c->cs), cx, cy, bw, bh;
    if (!(conf >= yolo->ignore_thres)) return 0;
    cx = (col + logistic(c->tx)) * netw / gw;
    cy = (row + logistic(c->ty)) * neth / gh;
    bw = (float)exp(c->tw) * yolo->anchor_list[anchor][0] * yolo->scale_x_y;
    bh = (float)exp(c->th) * yolo->anchor_list[anchor][1] * yolo->scale_x_y;
    out->type  = c->cls;
    out->score = conf;
    out->x1 = cx - bw * 0.5f; out->y1 = cy - bh * 0.5f;
    out->x2 = cx + bw * 0.5f; out->y2 = cy + bh * 0.5f;
    return 1;
}

/* Whole-head decode from a CHW head tensor (the single-image net_forward path). Returns new count. */
int ffb_decode_head_chw(const LAYER *yolo, const float *head, int gw, int gh, int netw, int neth,
                        BBOX *boxes, int n, int cap)
{
    const size_t plane = (size_t)gw * gh; const int per = 5 + yolo->class_num;
    int cell, a, l;
    for (cell = 0; cell < gw * gh; cell++) {
        for (a = 0; a < 3; a++) {
            const float *v = head + (size_t)a * per * plane + cell;
            ffb_candidate c; BBOX b;
            c.bs = v[4 * plane]; c.cs = v[5 * plane]; c.cls = 0;
            for (l = 1; l < yolo->class_num; l++) {
                float s = v[(size_t)(5 + l) * plane];
                if (c.cs < s) { c.cs = s; c.cls = l; }
            }
            c.tx = v[0]; c.ty = v[plane]; c.tw = v[2 * plane]; c.th = v[3 * plane];
            if (ffb_decode_candidate(yolo, netw, neth, gw, gh, cell, a, &c, &b) && n < cap) boxes[n++] = b;
        }
    }
    return n;
}

static int by_score_desc(const void *a, const void *b)
{
    const float x = ((const BBOX *)a)->score, y = ((const BBOX *)b)->score;
    return (x < y) - (x > y);
}

static float overlap_ratio(const BBOX *p, const BBOX *q, int min_mode)
{
    float left = p->x1 > q->x1 ? p->x1 : q->x1, top    = p->y1 > q->y1 ? p->y1 : q->y1;
    float right = p->x2 < q->x2 ? p->x2 : q->x2, bottom = p->y2 < q->y2 ? p->y2 : q->y2;
    float inter = (left < right && top < bottom) ? (right - left) * (bottom - top) : 0;
    float ap = (p->x2 - p->x1) * (p->y2 - p->y1), aq = (q->x2 - q->x1) * (q->y2 - q->y1);
    float uni = ap + aq - inter;
    return min_mode ? inter / (ap < aq ? ap : aq) : inter / uni;
}

int ffb_nms(BBOX *bx, int n, float threshold, int min_mode, int s1, int s2)
{
    int pivot, j, kept = 0;
    if (!bx || n <= 0) return 0;
    qsort(bx, (size_t)n, sizeof(BBOX), by_score_desc);
    /* pivot = current best surviving box; every later same-class box overlapping it is dropped
       (score := 0).  Next pivot = first later box that was looked at and survived. */
    for (pivot = 0; pivot >= 0 && pivot < n; ) {
        int next = -1;
        for (j = pivot + 1; j < n; j++) {
            if (bx[j].score == 0) continue;
            if (bx[j].type == bx[pivot].type && overlap_ratio(&bx[pivot], &bx[j], min_mode) > threshold) bx[j].score = 0;
            else if (next < 0) next = j;
        }
        pivot = next;
    }
    for (j = 0; j < n; j++) {
        BBOX b = bx[j];
        if (b.score == 0) continue;
        b.x1 = b.x1 * s1 / s2; b.y1 = b.y1 * s1 / s2;
        b.x2 = b.x2 * s1 / s2; b.y2 = b.y2 * s1 / s2;
        bx[kept++] = b;
    }
    memset(bx + kept, 0, sizeof(BBOX) * (size_t)(n - kept));
    return kept;
}

void ffb_fit_geometry(int w, int h, int W, int H, int *sw, int *sh, int *s1, int *s2)
{
    if (w * H > h * W) { *sw = W; *sh = W * h / w; *s1 = w; *s2 = *sw; }   /* fit width  */
    else               { *sh = H; *sw = H * w / h; *s1 = h; *s2 = *sh; }   /* fit height */
}
