/*
 * ffcnn_cli -- command-line detector on libffcnn_b200.so.
 *
 * Two modes:
 *   ffcnn_cli [n [file.bmp [cfg [weights]]]]
 *       the reference's test driver (ffcnn.c:552-593), same arguments, same output lines, same out.bmp:
 *       load the picture, net_load at the picture's size, n x (net_input + net_forward), print the timing,
 *       net_profile, one "score/category/rect" line per box, draw the boxes in green, save out.bmp.
 *       Uses only the reference API of include/ffcnn.h + include/bmpfile.h.
 *   ffcnn_cli --batch cfg weights a.bmp b.bmp ...   [--out prefix]
 *       the batched path of include/ffcnn_b200.h: every picture of the list (all must share one size) becomes one
 *       frame of a batch; ffb_detect_batch_u8 = one H2D copy, one graph replay, one candidate read-back.
 *       Prints the reference's box lines under a "frame i: file" heading; with --out, writes <prefix><i>.bmp.
 *
 * Unlike the reference driver it checks net_load's result: without a usable CUDA device the library refuses to load
 * (there is no CPU fallback) and the tool exits 2 with the library's message.
 *
 * Build: gcc -O2 -I include tools/ffcnn_cli.c -L ffcnn_b200 -lffcnn_b200 -Wl,-rpath,$PWD/ffcnn_b200 -o ffcnn_cli
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "ffcnn.h"
#include "ffcnn_b200.h"
#include "bmpfile.h"

static float MEAN[3] = { 0.0f, 0.0f, 0.0f };
static float NORM[3] = { 1 / 255.f, 1 / 255.f, 1 / 255.f };

static int now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (int)(ts.tv_sec * 1000 + ts.tv_nsec / 1000000);
}

static void report_boxes(BMP *pic, const BBOX *box, int count)
{
    for (int i = 0; i < count; i++) {
        const int x1 = (int)box[i].x1, y1 = (int)box[i].y1, x2 = (int)box[i].x2, y2 = (int)box[i].y2;
        printf("score: %.2f, category: %2d, rect: (%3d %3d %3d %3d)\n", box[i].score, box[i].type, x1, y1, x2, y2);
        if (pic) bmp_rectangle(pic, x1, y1, x2, y2, 0, 255, 0);
    }
}

static int run_single(int argc, char **argv)
{
    int   n       = argc > 1 ? atoi(argv[1]) : 10;
    char *picture = argc > 2 ? argv[2] : "test.bmp";
    char *cfg     = argc > 3 ? argv[3] : "yolo-fastest-1.1.cfg";
    char *weights = argc > 4 ? argv[4] : "yolo-fastest-1.1.weights";
    BMP   pic     = {0};

    printf("file_bmp    : %s\n", picture);
    printf("file_cfg    : %s\n", cfg);
    printf("file_weights: %s\n", weights);
    if (bmp_load(&pic, picture) != 0) { printf("failed to load bmp file: %s !\n", picture); return -1; }

    NET *net = net_load(cfg, weights, pic.width, pic.height);
    if (!net) { fprintf(stderr, "ffcnn_cli: net_load failed: %s\n", ffb_last_error()); bmp_free(&pic); return 2; }
    net_dump(net);

    const int t0 = now_ms();
    for (int i = 0; i < n; i++) {
        net_input(net, pic.pdata, pic.width, pic.height, MEAN, NORM);
        net_forward(net);
    }
    printf("%d times inference: %d ms\n", n, now_ms() - t0);
    net_profile(net);
    report_boxes(&pic, net->bbox_list, net->bbox_num);
    net_free(net);
    bmp_save(&pic, "out.bmp");
    bmp_free(&pic);
    return 0;
}

static int run_batch(int argc, char **argv)
{
    const char *prefix = NULL;
    char *files[4096];
    int   nfiles = 0;
    if (argc < 5) { fprintf(stderr, "usage: ffcnn_cli --batch cfg weights a.bmp [b.bmp ...] [--out prefix]\n"); return 1; }
    char *cfg = argv[2], *weights = argv[3];
    for (int i = 4; i < argc; i++) {
        if (!strcmp(argv[i], "--out") && i + 1 < argc) prefix = argv[++i];
        else if (nfiles < 4096) files[nfiles++] = argv[i];
    }
    if (nfiles == 0) { fprintf(stderr, "ffcnn_cli: no pictures\n"); return 1; }

    BMP *pics = calloc((size_t)nfiles, sizeof(BMP));
    int  rc = 0;
    for (int i = 0; i < nfiles && rc == 0; i++) {
        if (bmp_load(&pics[i], files[i]) != 0) { printf("failed to load bmp file: %s !\n", files[i]); rc = -1; }
        else if (pics[i].width != pics[0].width || pics[i].height != pics[0].height) {
            fprintf(stderr, "ffcnn_cli: %s is %dx%d, the batch is %dx%d\n", files[i], pics[i].width, pics[i].height, pics[0].width, pics[0].height);
            rc = 1;
        }
    }
    NET *net = NULL;
    unsigned char *frames = NULL;
    if (rc == 0) {
        const int w = pics[0].width, h = pics[0].height, pitch = pics[0].stride;
        const size_t frame_bytes = (size_t)h * pitch;
        net = ffb_net_parse(cfg, weights, w, h);
        if (!net || ffb_net_attach(net, getenv("FFCNN_DEVICE") ? atoi(getenv("FFCNN_DEVICE")) : 0, nfiles) != 0) {
            fprintf(stderr, "ffcnn_cli: %s\n", ffb_last_error()); rc = 2;
        }
        if (rc == 0 && !(frames = ffb_host_alloc_pinned(frame_bytes * nfiles))) { fprintf(stderr, "ffcnn_cli: %s\n", ffb_last_error()); rc = 2; }
        if (rc == 0) {
            for (int i = 0; i < nfiles; i++) memcpy(frames + i * frame_bytes, pics[i].pdata, frame_bytes);
            const int t0 = now_ms();
            if (ffb_detect_batch_u8(net, frames, nfiles, w, h, pitch, MEAN, NORM) != 0) { fprintf(stderr, "ffcnn_cli: %s\n", ffb_last_error()); rc = 2; }
            else printf("%d frames in one batch: %d ms\n", nfiles, now_ms() - t0);
        }
        for (int i = 0; i < nfiles && rc == 0; i++) {
            BBOX *box = NULL;
            const int count = ffb_boxes(net, i, &box);
            printf("frame %d: %s\n", i, files[i]);
            report_boxes(prefix ? &pics[i] : NULL, box, count < 0 ? 0 : count);
            if (prefix) {
                char name[1024];
                snprintf(name, sizeof name, "%s%d.bmp", prefix, i);
                bmp_save(&pics[i], name);
            }
        }
    }
    if (frames) ffb_host_free_pinned(frames);
    if (net) net_free(net);
    for (int i = 0; i < nfiles; i++) bmp_free(&pics[i]);
    free(pics);
    return rc;
}

int main(int argc, char **argv)
{
    if (argc > 1 && !strcmp(argv[1], "--batch")) return run_batch(argc, argv);
    return run_single(argc, argv);
}
