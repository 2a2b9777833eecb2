/*
 * oracle/ref_bench.c -- TEST/BENCH INFRASTRUCTURE (the "reference" CPU baseline).
 *
 * Links against the UNMODIFIED reference (ffcnn.c + conv-v6.c, build.sh flags,
 * compiled where they lie by oracle/Makefile) and times its own public API
 * (ffcnn.h:48-52) on the same synthetic workload bench.py gives the GPU path:
 * seeded 320x320x3 u8 BGR frames, net_load(cfg, weights, 0, 0).
 *
 * The reference is single-threaded, so host parallelism = P
 * independent processes (fork), each with its own NET; the parent reports the
 * aggregate frames/s.  usage:
 *   ffcnn_ref_bench <cfg> <weights> <procs> <frames_per_proc> [w h]
 * prints one line: "ref_bench procs=P frames=F seconds=S fps=X"
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <sys/wait.h>
#include <sys/mman.h>
#include "ffcnn.h"

static double now_s(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

/* same generator bench.py / tests use: splitmix64 stream, one byte per draw */
static uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static void fill_frame(unsigned char *p, int nbytes, uint64_t seed)
{
    uint64_t s = seed; int i;
    for (i = 0; i + 8 <= nbytes; i += 8) { uint64_t v = splitmix64(&s); memcpy(p + i, &v, 8); }
    if (i < nbytes) { uint64_t v = splitmix64(&s); memcpy(p + i, &v, nbytes - i); }
}

int main(int argc, char **argv)
{
    static float MEAN[3] = { 0, 0, 0 }, NORM[3] = { 1 / 255.f, 1 / 255.f, 1 / 255.f };
    int procs, frames, w = 320, h = 320, p, f;
    double *shared;
    if (argc < 5) { fprintf(stderr, "usage: %s cfg weights procs frames_per_proc [w h]\n", argv[0]); return 2; }
    procs = atoi(argv[3]); frames = atoi(argv[4]);
    if (argc > 6) { w = atoi(argv[5]); h = atoi(argv[6]); }
    if (procs < 1) procs = 1;
    shared = mmap(NULL, sizeof(double) * 2 * procs, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (shared == MAP_FAILED) return 3;

    for (p = 0; p < procs; p++) {
        pid_t pid = fork();
        if (pid == 0) {
            NET *net = net_load(argv[1], argv[2], w == 320 && h == 320 ? 0 : w, w == 320 && h == 320 ? 0 : h);
            int pitch = (w * 3 + 3) & ~3;
            unsigned char *img = malloc((size_t)pitch * h);
            double t0, t1; long nb = 0;
            if (!net || !img) _exit(4);
            fill_frame(img, pitch * h, 0xFFC0ull + p);
            net_input(net, img, w, h, MEAN, NORM); net_forward(net);            /* warm-up */
            t0 = now_s();
            for (f = 0; f < frames; f++) {
                fill_frame(img, 64, 0xFFC0ull + p * 100003ull + f);              /* perturb the frame a little */
                net_input(net, img, w, h, MEAN, NORM);
                net_forward(net);
                nb += net->bbox_num;
            }
            t1 = now_s();
            shared[2 * p] = t0; shared[2 * p + 1] = t1;
            (void)nb;
            net_free(net); free(img);
            _exit(0);
        } else if (pid < 0) return 5;
    }
    {
        int status, bad = 0; double tmin = 1e300, tmax = 0;
        while (wait(&status) > 0) if (!WIFEXITED(status) || WEXITSTATUS(status)) bad++;
        if (bad) { fprintf(stderr, "ref_bench: %d worker(s) failed\n", bad); return 6; }
        for (p = 0; p < procs; p++) { if (shared[2 * p] < tmin) tmin = shared[2 * p]; if (shared[2 * p + 1] > tmax) tmax = shared[2 * p + 1]; }
        printf("ref_bench procs=%d frames=%d seconds=%.6f fps=%.3f\n", procs, procs * frames, tmax - tmin, procs * frames / (tmax - tmin));
    }
    return 0;
}
