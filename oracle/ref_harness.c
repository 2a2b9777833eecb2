/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE, never part of the product path.
 *
 * Compiles the UNMODIFIED reference translation unit (ffcnn.c, pulled in from
 * /root/reference through the include path given by oracle/Makefile) together
 * with this file so the reference's file-static layer functions are callable,
 * and exposes three helpers the parity tests / golden generator need:
 *
 *   refh_layer_info   -- geometry of layer i as the reference parsed it
 *   refh_forward_dump -- same layer sequence as net_forward (ffcnn.c:476-520)
 *                        but keeps every layer output alive so it can be copied
 *                        out, and returns the raw pre-NMS candidates
 *   refh_groupconv    -- the reference operator seam (conv.h:4-7) of whichever
 *                        conv-vN.c this library was linked with
 *
 * Nothing of the reference is copied here: the sources are compiled where they
 * lie.  The output library lives in oracle/_ref/ (git-ignored).
 */
#include "ffcnn.c"

int refh_layer_num(NET *net) { return net ? net->layer_num : 0; }

/* info[0..11] = type,w,h,c (input of layer i), ow,oh,oc (its output), fs,stride,pad,groups,activation */
void refh_layer_info(NET *net, int i, int *info)
{
    LAYER *a = net->layer_list + i, *b = net->layer_list + i + 1;
    info[0] = a->type; info[1] = a->w; info[2] = a->h; info[3] = a->c;
    info[4] = b->w;    info[5] = b->h; info[6] = b->c;
    info[7] = a->fs;   info[8] = a->stride; info[9] = a->pad; info[10] = a->groups; info[11] = a->activation;
}

float *refh_input_ptr(NET *net) { return net->layer_list[0].data; }
float *refh_weight_buf(NET *net, int *nfloats) { *nfloats = net->weight_size; return net->weight_buf; }
int    refh_scale(NET *net, int *s1, int *s2) { *s1 = net->s1; *s2 = net->s2; return 0; }

/*
 * outs[i] (may be NULL) receives the output of layer i (CHW floats, ow*oh*oc).
 * raw receives up to raw_cap pre-NMS candidates; *nraw their count.
 * After return net->bbox_list / bbox_num hold the post-NMS result exactly as
 * net_forward would have left them.
 */
int refh_forward_dump(NET *net, float **outs, BBOX *raw, int raw_cap, int *nraw)
{
    int n = net->layer_num, i;
    /* bbox_list aliases the layer-0 input (ffcnn.c:243-244); the input is only read by
       layer 0 itself, before any yolo layer writes boxes, so the dump is unaffected. */
    for (i = 0; i < n; i++) {
        LAYER *il = net->layer_list + i, *ol = il + 1;
        size_t osz = (size_t)ol->w * ol->h * ol->c;
        if (il->type != LAYER_TYPE_DROPOUT && il->type != LAYER_TYPE_YOLO) {
            ol->data = malloc(osz * sizeof(float));
            if (!ol->data) return -1;
        }
        switch (il->type) {
        case LAYER_TYPE_CONV    : layer_groupconv_forward (net, il, ol); break;
        case LAYER_TYPE_AVGPOOL : layer_avgmaxpool_forward(il, ol, 0);   break;
        case LAYER_TYPE_MAXPOOL : layer_avgmaxpool_forward(il, ol, 1);   break;
        case LAYER_TYPE_UPSAMPLE: layer_upsample_forward  (il, ol);      break;
        case LAYER_TYPE_DROPOUT : ol->data = il->data;                   break; /* alias, keep il->data for the dump */
        case LAYER_TYPE_SHORTCUT: layer_shortcut_forward  (net, il, ol); break;
        case LAYER_TYPE_ROUTE   : layer_route_forward     (net, il, ol); break;
        case LAYER_TYPE_YOLO    : layer_yolo_forward      (net, il);     break;
        }
        if (outs && outs[i] && ol->data && il->type != LAYER_TYPE_YOLO) memcpy(outs[i], ol->data, osz * sizeof(float));
    }
    if (nraw) *nraw = net->bbox_num;
    if (raw)  memcpy(raw, net->bbox_list, sizeof(BBOX) * (net->bbox_num < raw_cap ? net->bbox_num : raw_cap));
    net->bbox_num = nms(net->bbox_list, net->bbox_num, 0.5f, 1, net->s1, net->s2);

    /* release everything we kept alive (dropout outputs alias their inputs) */
    for (i = n; i >= 1; i--) {
        LAYER *il = net->layer_list + i - 1, *ol = net->layer_list + i;
        if (il->type == LAYER_TYPE_DROPOUT) { ol->data = NULL; continue; }
        free(ol->data); ol->data = NULL;
    }
    return 0;
}

void refh_groupconv(float *datai, float *dataf, float *datao,
                    int iw, int ih, int ic, int ig, int ipad, int istride,
                    int fs, int fn, int ow, int oh, int oc, int activation)
{
    float *buf = NULL; int bufsize = 0;
    groupconv(datai, dataf, datao, iw, ih, ic, ig, ipad, istride, fs, fn, ow, oh, oc, activation, &buf, &bufsize);
    free(buf);
}

int refh_sizeof_bbox(void) { return (int)sizeof(BBOX); }
