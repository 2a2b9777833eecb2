/*
 * bmpfile.h -- 24-bit BMP helper with the reference's interface (bmpfile.h:9-22), host C.
 * Pixels are stored top-down, B,G,R per pixel, row pitch ALIGN(3*width, 4): exactly the
 * buffer net_input() expects (ffcnn.c:274).
 */
#ifndef FFCNN_B200_BMPFILE_H
#define FFCNN_B200_BMPFILE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int   width;
    int   height;
    int   stride;   /* bytes per row */
    int   cdepth;   /* bits per pixel (24) */
    void *pdata;
} BMP;

int  bmp_load     (BMP *pb, char *file);
int  bmp_save     (BMP *pb, char *file);
void bmp_free     (BMP *pb);
void bmp_setpixel (BMP *pb, int x, int y, int  r, int  g, int  b);
void bmp_getpixel (BMP *pb, int x, int y, int *r, int *g, int *b);
void bmp_rectangle(BMP *pb, int x1, int y1, int x2, int y2, int r, int g, int b);

#ifdef __cplusplus
}
#endif
#endif
