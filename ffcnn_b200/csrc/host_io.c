/*
 * host_io.c -- the thin host-C parts of the ffcnn.h / bmpfile.h boundary that are not GPU work.
 *
 *   net_input  (single frame) ... ffcnn.c:259-289: fills the host CHW tensor layer_list[0].data, which
 *                                 net_forward then uploads; the batched path (ffb_input_u8) does the
 *                                 same arithmetic in a CUDA kernel instead.
 *   net_dump / net_profile ...... ffcnn.c:522-550 (same table columns / per-type ms rows)
 *   net_free .................... ffcnn.c:249-257
 *   bmp_* ....................... bmpfile.c:42-156 (24-bit only, rows flipped to top-down on load)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ffb_internal.h"
#include "../../include/bmpfile.h"

void net_input(NET *net, unsigned char *bgr, int w, int h, float *mean, float *norm)
{
    ffb_net *fn; LAYER *l0; int sw, sh, s1, s2, i, j, pitch; size_t plane;
    if (!net || !bgr) return;
    fn = ffb_from_pub(net); l0 = net->layer_list;
    memset(net->bbox_list, 0, sizeof(BBOX) * (size_t)net->bbox_num);
    net->bbox_num = 0;
    ffb_fit_geometry(w, h, l0->w, l0->h, &sw, &sh, &s1, &s2);
    net->s1 = s1; net->s2 = s2;
    for (i = 0; i < 3; i++) { fn->in_mean[i] = mean[i]; fn->in_norm[i] = norm[i]; }
    pitch = FFB_ALIGN(w * 3, 4);
    plane = (size_t)l0->w * l0->h;
    for (i = 0; i < sh; i++) {
        const unsigned char *src = bgr + (size_t)(i * s1 / s2) * pitch;
        float *r = l0->data + (size_t)i * l0->w, *g = r + plane, *b = g + plane;
        for (j = 0; j < sw; j++) {
            const unsigned char *px = src + (j * s1 / s2) * 3;
            r[j] = (px[2] - mean[0]) * norm[0];
            g[j] = (px[1] - mean[1]) * norm[1];
            b[j] = (px[0] - mean[2]) * norm[2];
        }
    }
}

void net_forward(NET *net)
{
    ffb_net *fn;
    if (!net) return;
    fn = ffb_from_pub(net);
    if (fn->magic != FFB_MAGIC || !fn->engine) {
        fprintf(stderr, "ffcnn_b200: net_forward on a net without a GPU engine (no CPU fallback)\n");
        return;
    }
    if (ffb_engine_forward_single(fn) != 0) fprintf(stderr, "ffcnn_b200: net_forward failed: %s\n", ffb_last_error());
}

void net_free(NET *net)
{
    ffb_net *fn;
    if (!net) return;
    fn = ffb_from_pub(net);
    if (fn->engine) ffb_engine_destroy(fn->engine);
    free(net->layer_list[0].data);
    free(net->bbox_list);
    free(net->cnntempbuf);
    free(net->weight_buf);
    free(fn);
}

static const char *layer_name(int t)
{
    static const char *names[] = { "conv", "avgpool", "maxpool", "upsample", "dropout", "shortcut", "route", "yolo" };
    return t >= 0 && t < 8 ? names[t] : "unknown";
}

static const char *act_name(int a)
{
    return a == FFB_ACT_LINEAR ? "linear" : a == FFB_ACT_RELU ? "relu" : a == FFB_ACT_LEAKY ? "leaky" : "unknown";
}

void net_dump(NET *net)
{
    int i, j;
    if (!net) return;
    printf("layer   type  filters fltsize  pad/strd input          output       bn/act\n");
    for (i = 0; i < net->layer_num; i++) {
        const LAYER *a = net->layer_list + i, *b = a + 1;
        switch (a->type) {
        case LAYER_TYPE_YOLO:
            printf("%3d %8s class_num: %d ignore_thres: %3.2f [%d, %d] [%d, %d] [%d, %d]\n", i, layer_name(a->type),
                   a->class_num, a->ignore_thres, a->anchor_list[0][0], a->anchor_list[0][1],
                   a->anchor_list[1][0], a->anchor_list[1][1], a->anchor_list[2][0], a->anchor_list[2][1]);
            break;
        case LAYER_TYPE_DROPOUT:
            printf("%3d %8s %-38s -> %3dx%3dx%3d\n", i, layer_name(a->type), "", b->w, b->h, b->c);
            break;
        case LAYER_TYPE_SHORTCUT: case LAYER_TYPE_ROUTE: {
            char deps[256]; int n = snprintf(deps, sizeof deps, "layers:");
            for (j = 0; j < a->depend_num && n < (int)sizeof deps; j++) n += snprintf(deps + n, sizeof deps - (size_t)n, " %d", a->depend_list[j]);
            printf("%3d %8s %-38s -> %3dx%3dx%3d\n", i, layer_name(a->type), deps, b->w, b->h, b->c);
            break; }
        default:
            printf("%3d %8s %3d/%3d %2dx%2dx%3d   %d/%2d   %3dx%3dx%3d -> %3dx%3dx%3d  %d/%-6s\n", i, layer_name(a->type),
                   a->fn, a->groups, a->fs, a->fs, a->c / a->groups, a->pad, a->stride, a->w, a->h, a->c,
                   b->w, b->h, b->c, a->batchnorm, act_name(a->activation));
        }
    }
}

void net_profile(NET *net)
{
    int t;
    if (!net) return;
    for (t = 0; t < LAYER_TYPE_TOTOAL; t++) printf("%8s: %5d ms\n", layer_name(t), net->timeused[t]);
}

/* ------------------------------------------------------------------------------- BMP (24 bpp) */

#pragma pack(push, 1)
typedef struct {
    uint16_t magic; uint32_t file_size; uint16_t rsv1, rsv2; uint32_t data_offset;
    uint32_t info_size, width, height; uint16_t planes, bpp;
    uint32_t compression, image_size, xppm, yppm, clr_used, clr_important;
} bmp_header;
#pragma pack(pop)

int bmp_load(BMP *pb, char *file)
{
    bmp_header hd; FILE *f = fopen(file, "rb"); int y;
    if (!f) return -1;
    memset(&hd, 0, sizeof hd);
    if (fread(&hd, sizeof hd, 1, f) != 1) { fclose(f); return -1; }
    pb->width = (int)hd.width; pb->height = (int)hd.height;
    pb->stride = FFB_ALIGN(pb->width * 3, 4); pb->cdepth = 24;
    pb->pdata = malloc((size_t)pb->stride * pb->height);
    if (pb->pdata)                                     /* file rows are bottom-up */
        for (y = pb->height - 1; y >= 0; y--)
            if (fread((unsigned char *)pb->pdata + (size_t)y * pb->stride, (size_t)pb->stride, 1, f) != 1) break;
    fclose(f);
    return pb->pdata ? 0 : -1;
}

int bmp_save(BMP *pb, char *file)
{
    bmp_header hd; FILE *f; int y;
    memset(&hd, 0, sizeof hd);
    hd.magic = 0x4D42; hd.data_offset = sizeof hd; hd.info_size = 40;
    hd.width = (uint32_t)pb->width; hd.height = (uint32_t)pb->height; hd.planes = 1; hd.bpp = (uint16_t)pb->cdepth;
    hd.image_size = (uint32_t)(pb->stride * pb->height); hd.file_size = hd.data_offset + hd.image_size;
    f = fopen(file, "wb");
    if (!f) return -1;
    fwrite(&hd, sizeof hd, 1, f);
    for (y = pb->height - 1; y >= 0; y--) fwrite((unsigned char *)pb->pdata + (size_t)y * pb->stride, (size_t)pb->stride, 1, f);
    fclose(f);
    return 0;
}

void bmp_free(BMP *pb)
{
    free(pb->pdata);
    memset(pb, 0, sizeof *pb);
}

static int clamp255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

void bmp_setpixel(BMP *pb, int x, int y, int r, int g, int b)
{
    unsigned char *px;
    if (x < 0 || y < 0 || x >= pb->width || y >= pb->height) return;
    px = (unsigned char *)pb->pdata + (size_t)y * pb->stride + (size_t)x * (pb->cdepth / 8);
    px[0] = (unsigned char)clamp255(b); px[1] = (unsigned char)clamp255(g); px[2] = (unsigned char)clamp255(r);
}

void bmp_getpixel(BMP *pb, int x, int y, int *r, int *g, int *b)
{
    const unsigned char *px;
    if (x < 0 || y < 0 || x >= pb->width || y >= pb->height) { *r = *g = *b = 0; return; }
    px = (const unsigned char *)pb->pdata + (size_t)y * pb->stride + (size_t)x * (pb->cdepth / 8);
    *r = px[0]; *g = px[1]; *b = px[2];               /* channel order as the reference returns it (bmpfile.c:139-141) */
}

void bmp_rectangle(BMP *pb, int x1, int y1, int x2, int y2, int r, int g, int b)
{
    int i;
    for (i = x1; i <= x2; i++) { bmp_setpixel(pb, i, y1, r, g, b); bmp_setpixel(pb, i, y2, r, g, b); }
    for (i = y1; i <= y2; i++) { bmp_setpixel(pb, x1, i, r, g, b); bmp_setpixel(pb, x2, i, r, g, b); }
}
