/* auction_sim.c — CPU model of K4's phase 1 (epsilon = 0 auction over certified candidate lists, FIFO service,
 * W rows bidding per micro-round from the same price snapshot).  Developer tool: it produced the bid / refresh /
 * sequential-step counts quoted in DESIGN.md (FIFO vs depth-first service, list length vs refreshes, problems
 * without slack columns, the effect of the bid budget).  Not part of the product or of the test oracle.
 *
 *   gcc -O2 -o /tmp/auction_sim tools/auction_sim.c -lm
 *   /tmp/auction_sim cost.npy NR NC [K=128] [W=32] [T0=32] [M0=4] [RT=32] [RM=4] [BUDGET_PER_ROW=0]
 *
 * cost.npy: float32 C-order matrix written by numpy.save (128-byte header).  T lanes keep the M smallest reduced
 * values each (list build: T0/M0, refresh: RT/RM); tau = the smallest (M+1)-th value over the lanes.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

static int nr, nc, K, W, RT = 32, RM = 4;
static float *c;
static double *v, *tau;
static int *lcol, *llen;
static float *lcost;

static void build(int i, int T, int M) {
    const float *ci = c + (size_t)i * nc;
    static double best[1024][9];
    static int bj[1024][9];
    double t = INFINITY;
    for (int th = 0; th < T; th++) {
        for (int m = 0; m <= M; m++) { best[th][m] = INFINITY; bj[th][m] = -1; }
        for (int j = th; j < nc; j += T) {
            const double w = ci[j] - v[j];
            if (w < best[th][M]) {
                int m = M;
                best[th][M] = w; bj[th][M] = j;
                while (m > 0 && best[th][m] < best[th][m - 1]) {
                    double x = best[th][m]; best[th][m] = best[th][m - 1]; best[th][m - 1] = x;
                    int y = bj[th][m]; bj[th][m] = bj[th][m - 1]; bj[th][m - 1] = y;
                    m--;
                }
            }
        }
        if (best[th][M] < t) t = best[th][M];
    }
    int n = 0;
    for (int th = 0; th < T; th++)
        for (int m = 0; m < M; m++)
            if (bj[th][m] >= 0 && best[th][m] <= t) {
                if (n < K) { lcol[(size_t)i * K + n] = bj[th][m]; lcost[(size_t)i * K + n] = ci[bj[th][m]]; n++; }
                else if (best[th][m] < t) t = best[th][m];       /* what does not fit lowers tau */
            }
    llen[i] = n; tau[i] = t;
}

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s cost.npy NR NC [K W T0 M0 RT RM BUDGET]\n", argv[0]); return 2; }
    nr = atoi(argv[2]); nc = atoi(argv[3]);
    K = argc > 4 ? atoi(argv[4]) : 128; W = argc > 5 ? atoi(argv[5]) : 32;
    const int T0 = argc > 6 ? atoi(argv[6]) : 32, M0 = argc > 7 ? atoi(argv[7]) : 4;
    if (argc > 9) { RT = atoi(argv[8]); RM = atoi(argv[9]); }
    const long budget = argc > 10 ? atol(argv[10]) : 0;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    fseek(f, 128, SEEK_SET);
    c = malloc(sizeof(float) * (size_t)nr * nc);
    if (fread(c, 4, (size_t)nr * nc, f) != (size_t)nr * nc) { fprintf(stderr, "short read\n"); return 1; }
    fclose(f);
    v = calloc(nc, 8);
    int *r4c = malloc(4 * nc), *c4r = malloc(4 * nr);
    for (int j = 0; j < nc; j++) r4c[j] = -1;
    for (int i = 0; i < nr; i++) c4r[i] = -1;
    lcol = malloc(4 * (size_t)nr * K); lcost = malloc(4 * (size_t)nr * K); llen = malloc(4 * nr); tau = malloc(8 * nr);
    double avglen = 0;
    for (int i = 0; i < nr; i++) { build(i, T0, M0); avglen += llen[i]; }
    int qcap = 1;
    while (qcap < nr + 1) qcap <<= 1;
    int *q = malloc(4 * qcap), head = 0, count = nr;
    for (int i = 0; i < nr; i++) q[i] = i;
    long bids = 0, micro = 0, refresh = 0, parked = 0, lost = 0, short_rounds = 0;
    int row[64], col[64], kind[64];
    double gam[64];
    while (count > 0 && (!budget || bids < budget * nr)) {
        const int take = count < W ? count : W;
        micro++;
        if (take < W) short_rounds++;
        for (int w = 0; w < take; w++) {
            const int i = q[(head + w) & (qcap - 1)];
            double w1 = INFINITY, w2 = INFINITY;
            int j1 = -1;
            row[w] = i;
            for (int k = 0; k < llen[i]; k++) {
                const int j = lcol[(size_t)i * K + k];
                const double x = lcost[(size_t)i * K + k] - v[j];
                if (x < w1 || (x == w1 && j < j1)) { w2 = w1; w1 = x; j1 = j; } else if (x < w2) w2 = x;
            }
            if (!(w1 <= tau[i])) { kind[w] = 2; continue; }               /* list exhausted */
            double g = fmin(w2, tau[i]) - w1;
            if (!(g > 0)) g = 0;
            gam[w] = g; col[w] = j1;
            kind[w] = (g == 0 && r4c[j1] >= 0) ? 3 : 1;                     /* zero-increment steal: park */
        }
        head = (head + take) & (qcap - 1); count -= take;
        int push[128], np = 0;
        for (int w = 0; w < take; w++) {
            if (kind[w] == 1) {
                int beaten = 0;
                bids++;
                for (int o = 0; o < take; o++)
                    if (o != w && kind[o] == 1 && col[o] == col[w] && (gam[o] > gam[w] || (gam[o] == gam[w] && row[o] < row[w]))) beaten = 1;
                if (beaten) { push[np++] = row[w]; lost++; }
                else {
                    const int j = col[w], prev = r4c[j];
                    v[j] -= gam[w]; r4c[j] = row[w]; c4r[row[w]] = j;
                    if (prev >= 0) { c4r[prev] = -1; push[np++] = prev; }
                }
            } else if (kind[w] == 2) { build(row[w], RT, RM); refresh++; push[np++] = row[w]; }
            else parked++;
        }
        for (int k = 0; k < np; k++) { q[(head + count) & (qcap - 1)] = push[k]; count++; }
    }
    int asg = 0;
    for (int k = 0; k < nr; k++) asg += c4r[k] >= 0;
    printf("%d x %d  K=%d W=%d  initial list %.1f entries: micro-rounds %ld (not full: %ld)  bids %ld  lost %ld  refreshes %ld  "
           "parked %ld  assigned %d/%d\n", nr, nc, K, W, avglen / nr, micro, short_rounds, bids, lost, refresh, parked, asg, nr);
    return 0;
}
