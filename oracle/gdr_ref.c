/*
 * CPU oracle (plain C, fp32) for the GDKVM memory hot path: LKVA readout + Gated Delta Rule.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into, loaded by, or called from gdkvm_b200/.
 * Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: /root/reference ships no model code, no golden vectors and no numerical
 * tests (README.md:1,36-38; website/e2e/smoke.spec.ts:4-80; .MISSING_LARGE_BLOBS:1).  This is a
 * restatement of the token-recurrent equations in BASELINE.json north_star / BASELINE.md section 2
 * (concept named at README.md:20 and website/src/content/homepage/en.json:20), identical in
 * meaning to oracle/gdr_ref.py::gdr_recurrent_ref, threaded over the independent (clip, head)
 * chains with POSIX threads so the CPU baseline uses every host core.
 *
 *   per token i:  S <- exp(g_i) S ;  r = v_i - S^T k_i ;  S <- S + k_i (beta_i r)^T ;
 *                 o_i = scale * S^T q_i
 *
 * Layout: q,k [B,T,H,K]; v,o [B,T,H,V]; g,beta [B,T,H]; s0,sT [B,H,K,V]; all fp32, contiguous.
 */
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

int gdr_oracle_num_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    const float *q, *k, *v, *g, *beta, *s0;
    float *o, *sT;
    int B, T, H, K, V;
    float scale;
    int next_chain;           /* work queue: chains are claimed one at a time */
    int fail;
    pthread_mutex_t mu;
} gdr_job;

/* one independent (clip, head) chain, exactly the recurrence in the header */
static int run_chain(const gdr_job* J, int ch) {
    const int T = J->T, H = J->H, K = J->K, V = J->V;
    const int b = ch / H, h = ch % H;
    float* S = (float*)malloc((size_t)K * V * sizeof(float));
    float* r = (float*)malloc((size_t)V * sizeof(float));
    if (!S || !r) { free(S); free(r); return -1; }
    if (J->s0) memcpy(S, J->s0 + (size_t)ch * K * V, (size_t)K * V * sizeof(float));
    else memset(S, 0, (size_t)K * V * sizeof(float));
    for (int t = 0; t < T; ++t) {
        const size_t row = ((size_t)b * T + t) * H + h;
        const float* qi = J->q + row * K;
        const float* ki = J->k + row * K;
        const float* vi = J->v + row * V;
        float* oi = J->o + row * V;
        const float a = expf(J->g[row]);
        const float bt = J->beta[row];
        /* S <- a S ; r = v - S^T k */
        for (int x = 0; x < V; ++x) r[x] = 0.f;
        for (int d = 0; d < K; ++d) {
            float* Sd = S + (size_t)d * V;
            const float kd = ki[d];
            for (int x = 0; x < V; ++x) { Sd[x] *= a; r[x] += Sd[x] * kd; }
        }
        for (int x = 0; x < V; ++x) { r[x] = bt * (vi[x] - r[x]); oi[x] = 0.f; }
        /* S <- S + k (beta r)^T ; o = scale S^T q */
        for (int d = 0; d < K; ++d) {
            float* Sd = S + (size_t)d * V;
            const float kd = ki[d], qd = qi[d];
            for (int x = 0; x < V; ++x) { Sd[x] += kd * r[x]; oi[x] += Sd[x] * qd; }
        }
        for (int x = 0; x < V; ++x) oi[x] *= J->scale;
    }
    if (J->sT) memcpy(J->sT + (size_t)ch * K * V, S, (size_t)K * V * sizeof(float));
    free(S); free(r);
    return 0;
}

static void* worker(void* arg) {
    gdr_job* J = (gdr_job*)arg;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        const int ch = J->next_chain++;
        pthread_mutex_unlock(&J->mu);
        if (ch >= J->B * J->H) break;
        if (run_chain(J, ch) != 0) { pthread_mutex_lock(&J->mu); J->fail = 1; pthread_mutex_unlock(&J->mu); }
    }
    return NULL;
}

/* returns 0 on success, -1 on bad arguments / allocation failure; nthreads<=0 -> all online cores */
int gdr_oracle_recurrent_f32(const float* q, const float* k, const float* v, const float* g,
                             const float* beta, const float* s0 /* may be NULL */, float* o,
                             float* sT /* may be NULL */, int B, int T, int H, int K, int V,
                             float scale, int nthreads) {
    if (!q || !k || !v || !g || !beta || !o || B <= 0 || T < 0 || H <= 0 || K <= 0 || V <= 0) return -1;
    gdr_job J = {q, k, v, g, beta, s0, o, sT, B, T, H, K, V, scale, 0, 0, PTHREAD_MUTEX_INITIALIZER};
    if (nthreads <= 0) nthreads = gdr_oracle_num_threads();
    if (nthreads > B * H) nthreads = B * H;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    int started = 0;
    for (int i = 1; i < nthreads; ++i)
        if (pthread_create(&th[started], NULL, worker, &J) == 0) ++started;
    worker(&J);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    return J.fail ? -1 : 0;
}
