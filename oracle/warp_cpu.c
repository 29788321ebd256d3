/*
 * oracle/warp_cpu.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the RNN-T loss the reference calls at model.py:39,57,74
 * (`Warp_RNNTLoss(blank, reduction="mean")(logits, targets, act_lens, label_lens)`).
 * The arithmetic lives in a third-party module that is NOT in /root/reference:
 *   warp-transducer, fork YooSungHyun/warp-transducer, branch espnet_v1.1
 *   (reference README.md:9-10; no commit pin, absent from requirements.txt).
 * Its source is not on disk here, so this file restates the PUBLISHED algorithm
 * (Graves 2012, "Sequence Transduction with RNNs", eqs. 16-20, in the form
 * warp-transducer's CPU path uses it):
 *   1. log-softmax over V applied OUTSIDE the lattice code (the CPU wrapper does
 *      `log_softmax(acts, -1)` before calling cpu_rnnt);
 *   2. per utterance, independent of the others (`omp parallel for` over the
 *      minibatch): alpha forward sweep, beta backward sweep, gradient w.r.t. the
 *      log-probabilities at the blank and label positions;
 *   3. gradient w.r.t. the logits = log-softmax backward of (2)
 *      (g - softmax * sum_v g), which is what autograd does in the wrapper;
 *   4. cost_b = -log P(y|x) = -beta(0,0); reductions are done by the caller.
 * SURVEY.md section 8(a) "Exact maths to implement" is the same statement.
 *
 * PARITY PIN: the reference holds no tests or golden vectors for this path
 * ("parity unpinned" by the reference itself).  This oracle is pinned instead
 * against (i) KAT-1, the classic warp-transducer / torchaudio docstring vector
 * (cost 4.49566698 + full gradient) and (ii) torchaudio's CPU rnnt_loss -- the
 * loss the reference uses on its fp16 path (model.py:6,31) -- run in the build
 * container; see oracle/gen_golden.py and tests/test_oracle.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path never does.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define RNNT_ORACLE_OK 0
#define RNNT_ORACLE_INVALID 2

#define DEFINE_ORACLE(SUFFIX, REAL, EXP, LOG, LOG1P, FABS, NEG_INF)                               \
                                                                                                  \
    static inline REAL logaddexp_##SUFFIX(REAL a, REAL b) {                                       \
        if (a == NEG_INF) return b;                                                               \
        if (b == NEG_INF) return a;                                                               \
        REAL m = a > b ? a : b;                                                                   \
        return m + LOG1P(EXP(-FABS(a - b)));                                                      \
    }                                                                                             \
                                                                                                  \
    /* one utterance; lp = log-probs [T_max, U1_max, V] of this utterance (in/out scratch) */    \
    static void utterance_##SUFFIX(const REAL* logits, const int* labels, int T, int U, int Tm,   \
                                   int U1m, int V, int blank, REAL* lp, REAL* alpha, REAL* beta,  \
                                   REAL* cost, REAL* grad) {                                      \
        const int U1 = U + 1;                                                                     \
        (void)Tm;                                                                                 \
        /* 1. log-softmax over V for the valid box only */                                        \
        for (int t = 0; t < T; ++t)                                                               \
            for (int u = 0; u < U1; ++u) {                                                        \
                const REAL* x = logits + ((size_t)t * U1m + u) * V;                               \
                REAL* y = lp + ((size_t)t * U1m + u) * V;                                         \
                REAL m = x[0];                                                                    \
                for (int v = 1; v < V; ++v) m = x[v] > m ? x[v] : m;                              \
                REAL s = 0;                                                                       \
                for (int v = 0; v < V; ++v) s += EXP(x[v] - m);                                   \
                REAL lse = m + LOG(s);                                                            \
                for (int v = 0; v < V; ++v) y[v] = x[v] - lse;                                    \
            }                                                                                     \
        /* 2a. alpha */                                                                           \
        for (int t = 0; t < T; ++t)                                                               \
            for (int u = 0; u < U1; ++u) {                                                        \
                REAL a;                                                                           \
                if (t == 0 && u == 0) a = 0;                                                      \
                else {                                                                            \
                    REAL no_emit = NEG_INF, emit = NEG_INF;                                       \
                    if (t > 0)                                                                    \
                        no_emit = alpha[(size_t)(t - 1) * U1m + u] +                              \
                                  lp[((size_t)(t - 1) * U1m + u) * V + blank];                    \
                    if (u > 0)                                                                    \
                        emit = alpha[(size_t)t * U1m + u - 1] +                                   \
                               lp[((size_t)t * U1m + u - 1) * V + labels[u - 1]];                 \
                    a = logaddexp_##SUFFIX(no_emit, emit);                                        \
                }                                                                                 \
                alpha[(size_t)t * U1m + u] = a;                                                   \
            }                                                                                     \
        /* 2b. beta */                                                                            \
        for (int t = T - 1; t >= 0; --t)                                                          \
            for (int u = U1 - 1; u >= 0; --u) {                                                   \
                REAL b;                                                                           \
                if (t == T - 1 && u == U1 - 1) b = lp[((size_t)t * U1m + u) * V + blank];         \
                else {                                                                            \
                    REAL no_emit = NEG_INF, emit = NEG_INF;                                       \
                    if (t < T - 1)                                                                \
                        no_emit = beta[(size_t)(t + 1) * U1m + u] +                               \
                                  lp[((size_t)t * U1m + u) * V + blank];                          \
                    if (u < U1 - 1)                                                               \
                        emit = beta[(size_t)t * U1m + u + 1] +                                    \
                               lp[((size_t)t * U1m + u) * V + labels[u]];                         \
                    b = logaddexp_##SUFFIX(no_emit, emit);                                        \
                }                                                                                 \
                beta[(size_t)t * U1m + u] = b;                                                    \
            }                                                                                     \
        const REAL ll = beta[0];                                                                  \
        *cost = -ll;                                                                              \
        if (!grad) return;                                                                        \
        /* 2c + 3. d cost / d log-prob at blank/label, pushed through log-softmax backward */     \
        for (int t = 0; t < T; ++t)                                                               \
            for (int u = 0; u < U1; ++u) {                                                        \
                const size_t c = (size_t)t * U1m + u;                                             \
                const REAL* y = lp + c * V;                                                       \
                REAL* g = grad + c * V;                                                           \
                REAL g_blank = 0, g_label = 0;                                                    \
                if (t < T - 1)                                                                    \
                    g_blank = -EXP(alpha[c] + beta[c + U1m] + y[blank] - ll);                     \
                else if (u == U1 - 1)                                                             \
                    g_blank = -EXP(alpha[c] + y[blank] - ll);                                     \
                if (u < U1 - 1) g_label = -EXP(alpha[c] + beta[c + 1] + y[labels[u]] - ll);       \
                const REAL gsum = g_blank + g_label;                                              \
                for (int v = 0; v < V; ++v) g[v] = -EXP(y[v]) * gsum;                             \
                g[blank] += g_blank;                                                              \
                if (u < U1 - 1) g[labels[u]] += g_label;                                          \
            }                                                                                     \
    }                                                                                             \
                                                                                                  \
    int rnnt_oracle_cost_and_grad_##SUFFIX(const REAL* logits, const int* labels,                 \
                                           const int* act_lens, const int* label_lens, int B,     \
                                           int T, int U1, int V, int blank, REAL* costs,          \
                                           REAL* grads, REAL* alphas, REAL* betas,                \
                                           int num_threads) {                                     \
        if (B < 0 || T <= 0 || U1 <= 0 || V <= 0 || blank < 0 || blank >= V)                      \
            return RNNT_ORACLE_INVALID;                                                           \
        for (int b = 0; b < B; ++b) {                                                             \
            if (act_lens[b] <= 0 || act_lens[b] > T) return RNNT_ORACLE_INVALID;                  \
            if (label_lens[b] < 0 || label_lens[b] + 1 > U1) return RNNT_ORACLE_INVALID;          \
        }                                                                                         \
        const size_t cells = (size_t)T * U1;                                                      \
        if (grads) memset(grads, 0, sizeof(REAL) * cells * V * (size_t)B);                        \
        int bad = 0;                                                                              \
        (void)num_threads;                                                                        \
        _Pragma("omp parallel for schedule(dynamic) num_threads(num_threads > 0 ? num_threads : 1)") \
        for (int b = 0; b < B; ++b) {                                                             \
            REAL* lp = (REAL*)malloc(sizeof(REAL) * cells * V);                                   \
            REAL* a = alphas ? alphas + cells * b : (REAL*)malloc(sizeof(REAL) * cells);          \
            REAL* be = betas ? betas + cells * b : (REAL*)malloc(sizeof(REAL) * cells);           \
            if (!lp || !a || !be) { bad = 1; }                                                    \
            else {                                                                                \
                if (alphas) for (size_t i = 0; i < cells; ++i) a[i] = 0;                          \
                if (betas) for (size_t i = 0; i < cells; ++i) be[i] = 0;                          \
                utterance_##SUFFIX(logits + cells * V * b, labels + (size_t)(U1 - 1) * b,         \
                                   act_lens[b], label_lens[b], T, U1, V, blank, lp, a, be,        \
                                   costs + b, grads ? grads + cells * V * b : NULL);              \
            }                                                                                     \
            free(lp);                                                                             \
            if (!alphas) free(a);                                                                 \
            if (!betas) free(be);                                                                 \
        }                                                                                         \
        return bad ? 1 : RNNT_ORACLE_OK;                                                          \
    }

DEFINE_ORACLE(f32, float, expf, logf, log1pf, fabsf, (-INFINITY))
DEFINE_ORACLE(f64, double, exp, log, log1p, fabs, (-(double)INFINITY))

int rnnt_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
