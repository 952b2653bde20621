/*
 * rnnt_oracle.c -- plain-C restatement of the reference's transducer-loss algorithm.
 *
 * TEST ORACLE ONLY.  Nothing under tsasr_b200/ links, loads or calls this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may.
 *
 * Reference lines restated (paths relative to /root/reference, SB = vendor/speechbrain/speechbrain):
 *   - call site and integer length conversion ........ SB/nnet/losses.py:58-59,72-79
 *   - alpha recursion ................................ SB/nnet/loss/transducer_loss.py:60-106
 *   - beta recursion ................................. SB/nnet/loss/transducer_loss.py:139-180
 *   - gradient w.r.t. log-probs (Numba semantics) .... SB/nnet/loss/transducer_loss.py:200-236
 *   - loss / T_b and in-forward reduction ............ SB/nnet/loss/transducer_loss.py:104-106,280-287
 *   - torchaudio semantics (what the recipe runs): third-party torchaudio.functional.rnnt_loss,
 *     requirement "torchaudio>=0.9.0" (vendor/speechbrain/requirements.txt), installed build
 *     2.11.0+cu128, not vendored.  Published algorithm: Graves 2012 eq. 16-20 with the softmax
 *     Jacobian folded into the gradient:
 *        g[t,u,v] = p_v e^{a+b-L} - [v=blank] e^{a+lp_blank+b(t+1,u)-L} - [v=y_{u+1}] e^{a+lp_v+b(t,u+1)-L}
 *
 * Parity pinning: tests/test_oracle.py checks this file against the reference's known-answer test
 * (vendor/speechbrain/tests/unittests/test_losses.py:109-152), against tests/golden/ vectors
 * produced by the reference's Numba kernels and by torchaudio, and against torchaudio run live.
 *
 * Build: see oracle/Makefile (gcc -O2 -pthread -shared -fPIC; no OpenMP runtime in the image).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

/* Minimal pthread parallel-for over utterances (the image has no libgomp). */
typedef void (*utt_fn)(int b, void* ctx);
typedef struct { utt_fn fn; void* ctx; int n; volatile int next; } pf_t;
static int g_threads = 0;
static void* pf_worker(void* p) {
    pf_t* s = (pf_t*)p;
    for (;;) {
        int i = __sync_fetch_and_add(&s->next, 1);
        if (i >= s->n) break;
        s->fn(i, s->ctx);
    }
    return NULL;
}
static void parallel_for(int n, utt_fn fn, void* ctx) {
    int nt = g_threads > 0 ? g_threads : (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nt > n) nt = n;
    if (nt > 256) nt = 256;
    pf_t s = {fn, ctx, n, 0};
    if (nt <= 1) { pf_worker(&s); return; }
    pthread_t th[256];
    for (int i = 0; i < nt - 1; ++i) pthread_create(&th[i], NULL, pf_worker, &s);
    pf_worker(&s);
    for (int i = 0; i < nt - 1; ++i) pthread_join(th[i], NULL);
}

#define NEG_INF (-INFINITY)

/* max + log1p(exp(-|a-b|)) -- SB/nnet/loss/transducer_loss.py:94-96 */
static inline double logaddexp_d(double a, double b) {
    if (a == NEG_INF) return b;
    if (b == NEG_INF) return a;
    double m = a > b ? a : b;
    return m + log1p(exp(-fabs(a - b)));
}
static inline float logaddexp_f(float a, float b) {
    if (a == NEG_INF) return b;
    if (b == NEG_INF) return a;
    float m = a > b ? a : b;
    return m + log1pf(expf(-fabsf(a - b)));
}

/* Row log-sum-exp of V values (the "fused log-softmax" denominator). */
static double row_lse(const float* x, int V) {
    float m = x[0];
    for (int v = 1; v < V; ++v) if (x[v] > m) m = x[v];
    double s = 0.0;
    for (int v = 0; v < V; ++v) s += exp((double)x[v] - (double)m);
    return (double)m + log(s);
}

/*
 * One utterance.  skip/emit: [T,U] row-major views with leading stride U (lattice width).
 * alpha/beta: [Tb,Ub] scratch with stride U.  fp32 != 0 rounds every DP value to float, which is
 * what both reference implementations do.
 */
static double alpha_beta(const double* skip, const double* emit, int Tb, int Ub, int U,
                         double* alpha, double* beta, int fp32) {
    for (int t = 0; t < Tb; ++t) {
        for (int u = 0; u < Ub; ++u) {
            double v;
            if (t == 0 && u == 0) v = 0.0;
            else {
                double a = t > 0 ? alpha[(t - 1) * U + u] + skip[(t - 1) * U + u] : NEG_INF;
                double b = u > 0 ? alpha[t * U + u - 1] + emit[t * U + u - 1] : NEG_INF;
                v = fp32 ? (double)logaddexp_f((float)a, (float)b) : logaddexp_d(a, b);
            }
            alpha[t * U + u] = v;
        }
    }
    for (int t = Tb - 1; t >= 0; --t) {
        for (int u = Ub - 1; u >= 0; --u) {
            double v;
            if (t == Tb - 1 && u == Ub - 1) v = skip[t * U + u];
            else {
                double a = t < Tb - 1 ? beta[(t + 1) * U + u] + skip[t * U + u] : NEG_INF;
                double b = u < Ub - 1 ? beta[t * U + u + 1] + emit[t * U + u] : NEG_INF;
                v = fp32 ? (double)logaddexp_f((float)a, (float)b) : logaddexp_d(a, b);
            }
            beta[t * U + u] = v;
        }
    }
    return beta[0];
}

/*
 * torchaudio semantics.  logits [B,T,U,V] fp32 contiguous; targets [B,U-1] int32;
 * costs[B] = -log P; grads (may be NULL) [B,T,U,V] = d costs_b / d logits, zero outside Tb x Ub.
 * Returns 0, or a negative code for the precondition torchaudio also rejects.
 */
typedef struct {
    const float* logits; const int* targets; const int* logit_lengths; const int* target_lengths;
    int B, T, U, V, blank; float clamp; int fp32; float* costs; float* grads;
} ta_ctx;

static void ta_one(int b, void* p) {
    const ta_ctx* c = (const ta_ctx*)p;
    const int T = c->T, U = c->U, V = c->V, blank = c->blank, fp32 = c->fp32;
    const int Tb = c->logit_lengths[b], Ub = c->target_lengths[b] + 1;
    const int* tg = c->targets + (size_t)b * (U - 1);
    const size_t cells = (size_t)T * U;
    double* den = (double*)malloc(sizeof(double) * cells * 5);
    double *skip = den + cells, *emit = skip + cells, *alpha = emit + cells, *beta = alpha + cells;
    const float* lg = c->logits + (size_t)b * cells * V;
    for (int t = 0; t < Tb; ++t)
        for (int u = 0; u < Ub; ++u) {
            const float* row = lg + ((size_t)t * U + u) * V;
            const double d = row_lse(row, V);
            den[t * U + u] = d;
            skip[t * U + u] = (double)row[blank] - d;
            emit[t * U + u] = u < Ub - 1 ? (double)row[tg[u]] - d : NEG_INF;
            if (fp32) {
                skip[t * U + u] = (double)(float)skip[t * U + u];
                emit[t * U + u] = (double)(float)emit[t * U + u];
            }
        }
    const double L = alpha_beta(skip, emit, Tb, Ub, U, alpha, beta, fp32);
    c->costs[b] = (float)(-L);
    if (c->grads) {
        float* gb = c->grads + (size_t)b * cells * V;
        const float clamp = c->clamp;
        for (int t = 0; t < Tb; ++t)
            for (int u = 0; u < Ub; ++u) {
                const size_t cc = (size_t)t * U + u;
                const float* row = lg + cc * V;
                float* g = gb + cc * V;
                const double a = alpha[cc], occ = exp(a + beta[cc] - L), d = den[cc];
                for (int v = 0; v < V; ++v) g[v] = (float)(exp((double)row[v] - d) * occ);
                double bt1 = NEG_INF;
                if (t < Tb - 1) bt1 = beta[cc + U];
                else if (u == Ub - 1) bt1 = 0.0;
                if (bt1 != NEG_INF) g[blank] -= (float)exp(a + skip[cc] + bt1 - L);
                if (u < Ub - 1) g[tg[u]] -= (float)exp(a + emit[cc] + beta[cc + 1] - L);
                if (clamp > 0.f)
                    for (int v = 0; v < V; ++v) g[v] = g[v] > clamp ? clamp : (g[v] < -clamp ? -clamp : g[v]);
            }
    }
    free(den);
}

int rnnt_oracle_torchaudio(const float* logits, const int* targets, const int* logit_lengths,
                           const int* target_lengths, int B, int T, int U, int V, int blank,
                           float clamp, int fp32, float* costs, float* grads) {
    if (blank < 0) blank += V;
    if (blank < 0 || blank >= V) return -1;
    int maxT = 0, maxL = 0;
    for (int b = 0; b < B; ++b) {
        if (logit_lengths[b] > maxT) maxT = logit_lengths[b];
        if (target_lengths[b] > maxL) maxL = target_lengths[b];
    }
    if (maxT != T) return -2;      /* "input length mismatch"  */
    if (maxL + 1 != U) return -3;  /* "output length mismatch" */
    if (grads) memset(grads, 0, sizeof(float) * (size_t)B * T * U * V);
    ta_ctx c = {logits, targets, logit_lengths, target_lengths, B, T, U, V, blank, clamp, fp32, costs, grads};
    parallel_for(B, ta_one, &c);
    return 0;
}

/*
 * Numba semantics (Transducer.apply).  log_probs [B,maxT,maxU,V] already log-softmaxed;
 * labels [B,maxU-1]; Tl[b] frames, Ul[b] label count.  per_utt[b] = -log P / T_b.
 * grads (may be NULL): gradient w.r.t. log_probs, non-zero only at blank and labels[b,u].
 */
typedef struct {
    const float* lp; const int* labels; const int* Tl; const int* Ul;
    int B, maxT, maxU, V, blank, fp32; float* per_utt; float* grads;
} nb_ctx;

static void nb_one(int b, void* p) {
    const nb_ctx* c = (const nb_ctx*)p;
    const int U = c->maxU, V = c->V, blank = c->blank;
    const int Tb = c->Tl[b], Ub = c->Ul[b] + 1;
    const int* lab = c->labels + (size_t)b * (U - 1);
    const size_t cells = (size_t)c->maxT * U;
    double* skip = (double*)malloc(sizeof(double) * cells * 4);
    double *emit = skip + cells, *alpha = emit + cells, *beta = alpha + cells;
    const float* lp = c->lp + (size_t)b * cells * V;
    for (int t = 0; t < Tb; ++t)
        for (int u = 0; u < Ub; ++u) {
            const float* row = lp + ((size_t)t * U + u) * V;
            skip[t * U + u] = row[blank];
            emit[t * U + u] = u < Ub - 1 ? row[lab[u]] : NEG_INF;
        }
    const double L = alpha_beta(skip, emit, Tb, Ub, U, alpha, beta, c->fp32);
    c->per_utt[b] = (float)(-L / Tb);
    if (c->grads) {
        float* gb = c->grads + (size_t)b * cells * V;
        for (int t = 0; t < Tb; ++t)
            for (int u = 0; u < Ub; ++u) {
                const size_t cc = (size_t)t * U + u;
                double bt1 = NEG_INF;
                if (t < Tb - 1) bt1 = beta[cc + U];
                else if (u == Ub - 1) bt1 = 0.0;
                if (bt1 != NEG_INF) gb[cc * V + blank] = (float)(-exp(alpha[cc] + skip[cc] + bt1 - L));
                if (u < Ub - 1) gb[cc * V + lab[u]] = (float)(-exp(alpha[cc] + emit[cc] + beta[cc + 1] - L));
            }
    }
    free(skip);
}

int rnnt_oracle_numba(const float* log_probs, const int* labels, const int* Tl, const int* Ul,
                      int B, int maxT, int maxU, int V, int blank, int fp32,
                      float* per_utt, float* grads) {
    if (grads) memset(grads, 0, sizeof(float) * (size_t)B * maxT * maxU * V);
    nb_ctx c = {log_probs, labels, Tl, Ul, B, maxT, maxU, V, blank, fp32, per_utt, grads};
    parallel_for(B, nb_one, &c);
    return 0;
}

void rnnt_oracle_set_threads(int n) { g_threads = n; }

int rnnt_oracle_num_threads(void) {
    return g_threads > 0 ? g_threads : (int)sysconf(_SC_NPROCESSORS_ONLN);
}
