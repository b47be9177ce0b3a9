/* harness.c — TEST INFRASTRUCTURE: lets the reference's own main() (pi_sph_fluid.c:475-704), linked against
 * libsphb200.so instead of its operator functions (:44-236 and :241-411 deleted), run a bounded number of iterations and
 * leave its state in a file.  oracle/Makefile target `refmain` streams the reference source through sed (never
 * copied into this repo): the operators removed, `#define REALTIME` removed (the source says so for benchmarking, :10),
 * every `while(1)` turned into `while(sphb_refmain_continue())`, calculate_accelerations renamed to the
 * recording wrapper below.  Everything else — scene construction (:484-540), kick and drift loops (:615-624,
 * :637-640), statistics (:656-691), the two pthreads — is the reference's code as it stands.
 *
 *   SPHB_REFMAIN_STEPS  iterations of the main loop (default 100)
 *   SPHB_REFMAIN_OUT    file: int n | struct particle fluid[n] | float du_dt[n] | float dv_dt[n]
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

#include "sph_b200.h"

static pthread_t g_main;
static long g_done, g_max = 100;
static struct particle *g_fluid;
static float *g_du, *g_dv;
static struct neighbors_context *g_ctx;
static int g_n;

__attribute__((constructor)) static void harness_init(void)
{
    g_main = pthread_self();
    const char *s = getenv("SPHB_REFMAIN_STEPS");
    if (s) g_max = atol(s);
}

/* remembers the arrays main() owns (:491-493), then the library's operator (:303) */
void sphb_refmain_accelerations(float *du, float *dv, struct particle *fluid, struct particle *boundary,
                                struct neighbors_context *ctx_fluid, struct neighbors_context *ctx_boundary, float gx, float gy)
{
    g_du = du; g_dv = dv; g_fluid = fluid; g_ctx = ctx_fluid;
    calculate_accelerations(du, dv, fluid, boundary, ctx_fluid, ctx_boundary, gx, gy);
}

void sphb_refmain_set_n(int n) { if (!g_n) g_n = n; }

/* the loop condition of :611 (and of the two I/O threads' loops, :451 and :469, which simply end) */
int sphb_refmain_continue(void)
{
    if (!pthread_equal(pthread_self(), g_main)) return 0;
    if (g_done++ < g_max) return 1;
    const char *path = getenv("SPHB_REFMAIN_OUT");
    if (path && g_fluid && g_n > 0) {
        FILE *fp = fopen(path, "wb");
        if (!fp) { perror(path); exit(2); }
        fwrite(&g_n, sizeof g_n, 1, fp);
        fwrite(g_fluid, 28, (size_t)g_n, fp);
        fwrite(g_du, sizeof(float), (size_t)g_n, fp);
        fwrite(g_dv, sizeof(float), (size_t)g_n, fp);
        fclose(fp);
    }
    printf("refmain: %ld iterations of the reference main loop on libsphb200, n_fluid = %d\n", g_done - 1, g_n);
    fflush(stdout);
    exit(0);
}
