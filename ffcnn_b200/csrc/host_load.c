/*
 * host_load.c -- host half of net_load: darknet .cfg -> LAYER table, .weights -> packed rows.
 *
 * Written from the format the reference accepts (citations: /root/reference):
 *   sections and keys .......... ffcnn.c:50-62,128-208  (keys located by FIRST SUBSTRING match
 *                                inside the section, value = text after '='/' ' up to end of line,
 *                                ffcnn.c:64-84; a missing key reads as "" -> 0)
 *   geometry .................... effective pad = pad ? size/2 : 0 (ffcnn.c:145);
 *                                conv out = (in - k + 2p)/s + 1 (148-149); pool out = in/stride
 *                                (156-157); upsample out = in*stride (162-163); inputw/h override
 *                                rounded up to a multiple of 32 (133-134)
 *   weights file ................ 20-byte header, then per conv layer: fn bias, [fn scale, fn mean,
 *                                fn var], fn*(c/g)*k*k weights (ffcnn.c:107-112,211-239; readme.txt:77-97)
 *   packed row .................. ALIGN(k*k*c/g,4) weights + {scale', bias', mean, var} with
 *                                scale' = scale / (float)sqrt(var + 1e-5f), bias' = bias - mean*scale'
 *                                (ffcnn.c:222-233); no BN: scale' = 1, bias' = bias
 * No GPU is touched here; ffb_net_attach (engine.cu) does the device half.
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ffb_internal.h"

static __thread char g_err[512];

void ffb_set_error(const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
const char *ffb_last_error(void) { return g_err; }

static char *slurp(const char *path, long *len)
{
    FILE *f = fopen(path, "rb"); char *buf; long n;
    if (!f) return NULL;
    fseek(f, 0, SEEK_END); n = ftell(f); fseek(f, 0, SEEK_SET);
    buf = (char *)malloc((size_t)n + 1);
    if (buf) { n = (long)fread(buf, 1, (size_t)n, f); buf[n] = 0; if (len) *len = n; }
    fclose(f);
    return buf;
}

typedef enum { SEC_NET, SEC_CONV, SEC_AVG, SEC_MAX, SEC_UPSAMPLE, SEC_DROPOUT, SEC_SHORTCUT, SEC_ROUTE, SEC_YOLO, SEC_OTHER } sec_kind;

static int starts(const char *s, const char *prefix) { return strncmp(s, prefix, strlen(prefix)) == 0; }

static sec_kind classify(const char *s)
{
    if (starts(s, "[net]")) return SEC_NET;
    if (starts(s, "[conv]") || starts(s, "[convolutional]")) return SEC_CONV;
    if (starts(s, "[avg]") || starts(s, "[avgpool]")) return SEC_AVG;
    if (starts(s, "[max]") || starts(s, "[maxpool]")) return SEC_MAX;
    if (starts(s, "[upsample]")) return SEC_UPSAMPLE;
    if (starts(s, "[dropout]")) return SEC_DROPOUT;
    if (starts(s, "[shortcut]")) return SEC_SHORTCUT;
    if (starts(s, "[route]")) return SEC_ROUTE;
    if (starts(s, "[yolo]")) return SEC_YOLO;
    return SEC_OTHER;
}

/* value of `key` inside [sec, end): first substring hit wins; end == NULL means "to end of text" */
static const char *value_of(const char *sec, const char *end, const char *key, char *out, size_t cap)
{
    const char *p = strstr(sec, key); size_t n = 0;
    out[0] = 0;
    if (!p || (end && p >= end)) return out;
    p += strlen(key);
    while (*p == '=' || *p == ' ') p++;
    while (*p && *p != '\n' && n + 1 < cap) out[n++] = *p++;
    out[n] = 0;
    return out;
}

static int int_of(const char *sec, const char *end, const char *key)
{
    char v[256]; return atoi(value_of(sec, end, key, v, sizeof v));
}

static int activation_code(const char *s)
{
    if (starts(s, "linear")) return FFB_ACT_LINEAR;
    if (starts(s, "relu"))   return FFB_ACT_RELU;
    if (starts(s, "leaky"))  return FFB_ACT_LEAKY;
    return -1;                                   /* behaves as linear (utils.h:22) */
}

/* comma separated ints, at most cap of them */
static int int_list(const char *s, int *out, int cap)
{
    int n = 0;
    while (*s && n < cap) {
        while (*s == ',') s++;
        if (!*s) break;
        out[n++] = atoi(s);
        while (*s && *s != ',') s++;
    }
    return n;
}

static int filter_row_floats(const LAYER *l) { return FFB_ALIGN(l->fs * l->fs * (l->c / l->groups), 4) + 4; }

static int read_floats(FILE *f, float *dst, int stride, int n)
{
    int i, got = 0;
    for (i = 0; i < n; i++) got += (int)fread(dst + (size_t)i * stride, sizeof(float), 1, f);
    return got;
}

static void load_weights(ffb_net *fn, const char *path)
{
    NET *net = &fn->pub; FILE *f = path ? fopen(path, "rb") : NULL; float *cursor = net->weight_buf; int i, j;
    if (f) fseek(f, 20, SEEK_SET);               /* {int32 major, minor, revision; uint64 seen} */
    for (i = 0; i < net->layer_num; i++) {
        LAYER *l = net->layer_list + i; int row, taps;
        if (l->type != LAYER_TYPE_CONV) continue;
        row = filter_row_floats(l); taps = l->fs * l->fs * (l->c / l->groups);
        l->filter = cursor; cursor += (size_t)l->fn * row;
        if (!f) continue;                        /* zero weights, as the reference */
        {
            float *scale = l->filter + row - 4, *bias = scale + 1, *mean = scale + 2, *var = scale + 3;
            for (j = 0; j < l->fn; j++) scale[(size_t)j * row] = 1.0f;
            read_floats(f, bias, row, l->fn);
            if (l->batchnorm) {
                read_floats(f, scale, row, l->fn);
                read_floats(f, mean,  row, l->fn);
                read_floats(f, var,   row, l->fn);
                for (j = 0; j < l->fn; j++) {
                    size_t o = (size_t)j * row;
                    scale[o] = scale[o] / (float)sqrt(var[o] + 0.00001f);
                    bias[o]  = bias[o] - mean[o] * scale[o];
                }
            }
            for (j = 0; j < l->fn; j++)
                if (fread(l->filter + (size_t)j * row, sizeof(float), (size_t)taps, f) != (size_t)taps) break;
        }
    }
    if (f) fclose(f);
}

NET *ffb_net_parse(const char *cfgfile, const char *weightsfile, int inputw, int inputh)
{
    char *text = cfgfile ? slurp(cfgfile, NULL) : NULL, v[256];
    const char *p; int nlayers = 0, cur = 0; ffb_net *fn; NET *net;
    if (!text) { ffb_set_error("cannot read cfg '%s'", cfgfile ? cfgfile : "(null)"); return NULL; }

    for (p = strchr(text, '['); p; p = strchr(p + 1, '[')) {
        sec_kind k = classify(p);
        if (k != SEC_NET && k != SEC_OTHER) nlayers++;
    }
    fn = (ffb_net *)calloc(1, sizeof(ffb_net) + (size_t)(nlayers + 1) * sizeof(LAYER));
    if (!fn) { free(text); ffb_set_error("out of memory"); return NULL; }
    fn->magic = FFB_MAGIC;
    net = &fn->pub;
    net->layer_list = (LAYER *)(fn + 1);
    net->layer_num  = nlayers;

    for (p = strchr(text, '['); p; ) {
        const char *next = strchr(p + 1, '[');
        const char *end  = next ? next - 1 : NULL;      /* the reference clips one char early (ffcnn.c:129) */
        LAYER *il = net->layer_list + cur, *ol = il + 1;
        sec_kind k = classify(p);
        int is_layer = 1, i;
        il->stride = il->groups = 1;
        switch (k) {
        case SEC_NET:
            net->layer_list[0].w = inputw ? FFB_ALIGN(inputw, 32) : int_of(p, end, "width");
            net->layer_list[0].h = inputh ? FFB_ALIGN(inputh, 32) : int_of(p, end, "height");
            net->layer_list[0].c = int_of(p, end, "channels");
            is_layer = 0;
            break;
        case SEC_CONV:
            il->type   = LAYER_TYPE_CONV;
            il->fn     = int_of(p, end, "filters");
            il->fs     = int_of(p, end, "size");
            if ((i = int_of(p, end, "stride"))) il->stride = i;
            if ((i = int_of(p, end, "groups"))) il->groups = i;
            il->pad    = int_of(p, end, "pad") ? il->fs / 2 : 0;
            il->batchnorm  = int_of(p, end, "batch_normalize") != 0;
            il->activation = activation_code(value_of(p, end, "activation", v, sizeof v));
            ol->c = il->fn;
            ol->w = (il->w - il->fs + 2 * il->pad) / il->stride + 1;
            ol->h = (il->h - il->fs + 2 * il->pad) / il->stride + 1;
            net->weight_size += il->fn * filter_row_floats(il);
            break;
        case SEC_AVG: case SEC_MAX:
            il->type = k == SEC_AVG ? LAYER_TYPE_AVGPOOL : LAYER_TYPE_MAXPOOL;
            il->fs   = int_of(p, end, "size");
            if ((i = int_of(p, end, "stride"))) il->stride = i;
            ol->c = il->c; ol->w = il->w / il->stride; ol->h = il->h / il->stride;
            break;
        case SEC_UPSAMPLE:
            il->type = LAYER_TYPE_UPSAMPLE;
            if ((i = int_of(p, end, "stride"))) il->stride = i;
            ol->c = il->c; ol->w = il->w * il->stride; ol->h = il->h * il->stride;
            break;
        case SEC_DROPOUT:
            il->type = LAYER_TYPE_DROPOUT;
            ol->c = il->c; ol->w = il->w; ol->h = il->h;
            break;
        case SEC_SHORTCUT:
            il->type = LAYER_TYPE_SHORTCUT;
            il->depend_list[0] = int_of(p, end, "from") + cur;
            il->depend_num     = 1;
            il->activation     = activation_code(value_of(p, end, "activation", v, sizeof v));
            ol->c = il->c; ol->w = il->w; ol->h = il->h;
            break;
        case SEC_ROUTE: {
            int deps[4], n = int_list(value_of(p, end, "layers", v, sizeof v), deps, 4);
            il->type = LAYER_TYPE_ROUTE;
            for (i = 0; i < n; i++) {
                int d = deps[i] > 0 ? deps[i] : cur + deps[i];
                if (d < 0 || d >= cur) { ffb_set_error("route layer %d depends on layer %d", cur, d); free(fn); free(text); return NULL; }
                il->depend_list[i] = d;
                ol->c += net->layer_list[d + 1].c;
                ol->w  = net->layer_list[d + 1].w;
                ol->h  = net->layer_list[d + 1].h;
            }
            il->depend_num = n;
            break; }
        case SEC_YOLO: {
            int mask[9] = {0}, anch[18] = {0};
            il->type       = LAYER_TYPE_YOLO;
            il->class_num  = int_of(p, end, "classes");
            value_of(p, end, "scale_x_y", v, sizeof v);
            il->scale_x_y  = v[0] ? (float)atof(v) : 1.0f;
            il->ignore_thres = (float)atof(value_of(p, end, "ignore_thresh", v, sizeof v));
            int_list(value_of(p, end, "mask", v, sizeof v), mask, 9);
            int_list(value_of(p, end, "anchors", v, sizeof v), anch, 18);
            for (i = 0; i < 3; i++) {
                int m = mask[i] >= 0 && mask[i] < 9 ? mask[i] : 0;
                il->anchor_list[i][0] = anch[2 * m]; il->anchor_list[i][1] = anch[2 * m + 1];
            }
            break; }
        default:
            is_layer = 0;
        }
        if (k == SEC_SHORTCUT && (il->depend_list[0] < 0 || il->depend_list[0] >= cur)) {
            ffb_set_error("shortcut layer %d depends on layer %d", cur, il->depend_list[0]); free(fn); free(text); return NULL;
        }
        if (is_layer) cur++;
        p = next;
    }
    free(text);

    net->weight_buf = (float *)calloc((size_t)(net->weight_size > 0 ? net->weight_size : 1), sizeof(float));
    {
        LAYER *l0 = net->layer_list; size_t in_floats = (size_t)l0->w * l0->h * l0->c;
        l0->data       = (float *)calloc(in_floats ? in_floats : 1, sizeof(float));
        net->bbox_max  = (int)(in_floats * sizeof(float) / sizeof(BBOX));      /* capacity rule of ffcnn.c:243 */
        net->bbox_list = (BBOX *)calloc((size_t)(net->bbox_max > 0 ? net->bbox_max : 1), sizeof(BBOX));
        fn->input_w = l0->w; fn->input_h = l0->h; fn->input_c = l0->c;
        if (!net->weight_buf || !l0->data || !net->bbox_list) {
            ffb_set_error("out of memory"); free(net->weight_buf); free(l0->data); free(net->bbox_list); free(fn); return NULL;
        }
    }
    net->s1 = net->s2 = 1;
    load_weights(fn, weightsfile);
    return net;
}
