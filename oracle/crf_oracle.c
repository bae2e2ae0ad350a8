/*
 * crf_oracle.c — CPU oracle for the GECCO ClusterCRF marginal-inference path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under gecco_b200/ may link, import or call this file; it is
 * used by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference leg
 * as the checker and as the CPU timing baseline ("port" of the reference's CPU path).
 *
 * What it restates (f64 throughout, like CRFsuite's floatval_t = double):
 *   - the arithmetic of sklearn_crfsuite.CRF.predict_marginals_single(), i.e. python-crfsuite
 *     Tagger.set() + Tagger.marginal() on CRFsuite 0.12 (third-party, NOT vendored in
 *     /root/reference; pinned by pyproject.toml:43 `sklearn-crfsuite ~=0.5.0`).  Published
 *     algorithm (crf1d_tag.c state scoring, crf1d_context.c alpha/beta/marginal), SURVEY.md
 *     Appendix B: state scores s_t[l] = sum over the item's attributes of W[a][l] (value 1.0,
 *     unknown attributes dropped), E = exp(s), M = exp(trans); scaled forward
 *     (c_t = 1/sum alpha_t, 1 when the sum is 0), backward started at c_{T-1}, marginal
 *     alpha_t[l]*beta_t[l]/c_t.
 *   - the window / pad / max-pool loop of gecco/crf/__init__.py:209-258 and
 *     gecco/_meta.py:124-132 (sliding_window): one "set + marginals" per W-gene window, nothing
 *     reused across windows, exactly as the reference drives the tagger.
 *
 * Parity pinning: checked against the reference's own golden vector
 * tests/test_cli/data/BGC0001866.genes.tsv (python-crfsuite output, 23 genes) in
 * tests/test_oracle.py via tests/golden/bgc0001866.json — max |diff| 5.6e-16.
 *
 * Build: see oracle/Makefile (gcc -O2 -shared -fPIC -pthread).
 */

#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_OK 0
#define ORACLE_EINVAL -1
#define ORACLE_ENOMEM -2

typedef struct {
    int32_t A, L;
    const double *state_w; /* [A][L] */
    double *exp_trans;     /* [L][L], exp of the transition weights, from -> to */
} oracle_model;

/* Scratch of one tagger instance: sized for the longest sequence it has seen. */
typedef struct {
    int32_t cap;
    double *state; /* [T][L] state scores, then their exponentials */
    double *alpha; /* [T][L] */
    double *beta;  /* [T][L] */
    double *scale; /* [T]    */
    double *row;   /* [L]    */
} oracle_ctx;

static int ctx_reserve(oracle_ctx *ctx, int32_t T, int32_t L) {
    if (T <= ctx->cap) return ORACLE_OK;
    free(ctx->state); free(ctx->alpha); free(ctx->beta); free(ctx->scale); free(ctx->row);
    ctx->state = (double *)malloc(sizeof(double) * (size_t)T * L);
    ctx->alpha = (double *)malloc(sizeof(double) * (size_t)T * L);
    ctx->beta = (double *)malloc(sizeof(double) * (size_t)T * L);
    ctx->scale = (double *)malloc(sizeof(double) * (size_t)T);
    ctx->row = (double *)malloc(sizeof(double) * (size_t)L);
    if (!ctx->state || !ctx->alpha || !ctx->beta || !ctx->scale || !ctx->row) return ORACLE_ENOMEM;
    ctx->cap = T;
    return ORACLE_OK;
}

static void ctx_release(oracle_ctx *ctx) {
    free(ctx->state); free(ctx->alpha); free(ctx->beta); free(ctx->scale); free(ctx->row);
    memset(ctx, 0, sizeof(*ctx));
}

/*
 * "Tagger.set": state scores of a T-item sequence.  Item t is either empty (item_gene[t] < 0, the
 * padding dict `{}` of gecco/crf/__init__.py:227) or the attribute list of gene item_gene[t].
 * Every listed attribute counts with value 1.0 (pycrfsuite: True -> 1.0); ids outside [0, A) are
 * attributes absent from the model dictionary and are dropped.
 */
static void score_items(const oracle_model *m, oracle_ctx *ctx, const int64_t *gene_ptr,
                        const int32_t *attr_idx, const int64_t *item_gene, int32_t T) {
    const int32_t L = m->L;
    memset(ctx->state, 0, sizeof(double) * (size_t)T * L);
    for (int32_t t = 0; t < T; ++t) {
        const int64_t g = item_gene[t];
        if (g < 0) continue;
        double *s = ctx->state + (size_t)t * L;
        for (int64_t p = gene_ptr[g]; p < gene_ptr[g + 1]; ++p) {
            const int32_t a = attr_idx[p];
            if (a < 0 || a >= m->A) continue;
            const double *w = m->state_w + (size_t)a * L;
            for (int32_t l = 0; l < L; ++l) s[l] += w[l];
        }
    }
}

/* exp of the state scores, scaled forward, backward.  After this, marginal(t,l) is O(1). */
static void forward_backward(const oracle_model *m, oracle_ctx *ctx, int32_t T) {
    const int32_t L = m->L;
    const double *M = m->exp_trans;
    for (size_t k = 0; k < (size_t)T * L; ++k) ctx->state[k] = exp(ctx->state[k]);

    /* forward: alpha_0 = E_0; alpha_t[j] = (sum_i alpha_{t-1}[i] M[i][j]) E_t[j]; each row rescaled */
    for (int32_t t = 0; t < T; ++t) {
        double *cur = ctx->alpha + (size_t)t * L;
        const double *e = ctx->state + (size_t)t * L;
        if (t == 0) {
            for (int32_t j = 0; j < L; ++j) cur[j] = e[j];
        } else {
            const double *prev = cur - L;
            for (int32_t j = 0; j < L; ++j) cur[j] = 0.0;
            for (int32_t i = 0; i < L; ++i)
                for (int32_t j = 0; j < L; ++j) cur[j] += prev[i] * M[(size_t)i * L + j];
            for (int32_t j = 0; j < L; ++j) cur[j] *= e[j];
        }
        double sum = 0.0;
        for (int32_t j = 0; j < L; ++j) sum += cur[j];
        const double c = (sum != 0.0) ? 1.0 / sum : 1.0;
        ctx->scale[t] = c;
        for (int32_t j = 0; j < L; ++j) cur[j] *= c;
    }

    /* backward: beta_{T-1} = c_{T-1}; beta_t[i] = c_t * sum_j M[i][j] E_{t+1}[j] beta_{t+1}[j] */
    {
        double *last = ctx->beta + (size_t)(T - 1) * L;
        for (int32_t i = 0; i < L; ++i) last[i] = ctx->scale[T - 1];
    }
    for (int32_t t = T - 2; t >= 0; --t) {
        double *cur = ctx->beta + (size_t)t * L;
        const double *next = cur + L;
        const double *e = ctx->state + (size_t)(t + 1) * L;
        for (int32_t j = 0; j < L; ++j) ctx->row[j] = next[j] * e[j];
        for (int32_t i = 0; i < L; ++i) {
            double dot = 0.0;
            for (int32_t j = 0; j < L; ++j) dot += M[(size_t)i * L + j] * ctx->row[j];
            cur[i] = dot * ctx->scale[t];
        }
    }
}

static inline double marginal(const oracle_ctx *ctx, int32_t L, int32_t t, int32_t l) {
    return ctx->alpha[(size_t)t * L + l] * ctx->beta[(size_t)t * L + l] / ctx->scale[t];
}

/* ------------------------------------------------------------------------------------------ */

static int model_init(oracle_model *m, const double *state_w, int32_t A, int32_t L, const double *trans_w) {
    if (!state_w || !trans_w || A < 0 || L <= 0) return ORACLE_EINVAL;
    m->A = A; m->L = L; m->state_w = state_w;
    m->exp_trans = (double *)malloc(sizeof(double) * (size_t)L * L);
    if (!m->exp_trans) return ORACLE_ENOMEM;
    for (int32_t k = 0; k < L * L; ++k) m->exp_trans[k] = exp(trans_w[k]);
    return ORACLE_OK;
}

/*
 * Primitive: marginals of every label at every item of ONE sequence given as CSR rows
 * [row_begin, row_end) — the equivalent of predict_marginals_single(xseq).  out is [T][L].
 */
int oracle_chain_marginals(const double *state_w, int32_t A, int32_t L, const double *trans_w,
                           const int64_t *gene_ptr, const int32_t *attr_idx, int64_t row_begin,
                           int64_t row_end, double *out) {
    oracle_model m; oracle_ctx ctx; memset(&ctx, 0, sizeof(ctx));
    const int64_t T64 = row_end - row_begin;
    if (T64 <= 0 || T64 > INT32_MAX || !out) return ORACLE_EINVAL;
    const int32_t T = (int32_t)T64;
    int rc = model_init(&m, state_w, A, L, trans_w);
    if (rc) return rc;
    rc = ctx_reserve(&ctx, T, L);
    if (rc == ORACLE_OK) {
        int64_t *items = (int64_t *)malloc(sizeof(int64_t) * (size_t)T);
        if (!items) rc = ORACLE_ENOMEM;
        else {
            for (int32_t t = 0; t < T; ++t) items[t] = row_begin + t;
            score_items(&m, &ctx, gene_ptr, attr_idx, items, T);
            forward_backward(&m, &ctx, T);
            for (int32_t t = 0; t < T; ++t)
                for (int32_t l = 0; l < L; ++l) out[(size_t)t * L + l] = marginal(&ctx, L, t, l);
            free(items);
        }
    }
    ctx_release(&ctx); free(m.exp_trans);
    return rc;
}

/*
 * One contig through the reference loop (gecco/crf/__init__.py:209-258, protein features):
 *   n < W, pad   -> delta = W-n; delta/2 empty items in front, (delta+1)/2 behind (:226-227)
 *   n < W, !pad  -> contig skipped, genes keep "no probability" (NaN here) (:228-234, :246-248)
 *   windows [i, i+W) for i in range(0, len+1-W, step) (_meta.py:131); probabilities start at 0
 *   and take the max over the windows covering them (:251-254); read back from delta/2 (:258).
 */
static int contig_windowed(const oracle_model *m, oracle_ctx *ctx, int64_t *items, double *prob,
                           const int64_t *gene_ptr, const int32_t *attr_idx, int64_t g0, int64_t n,
                           int32_t W, int32_t step, int32_t pad, int32_t pos_label, double *out,
                           int64_t *windows_done) {
    int64_t len = n, delta = 0;
    if (n < W) {
        if (!pad) {
            for (int64_t k = 0; k < n; ++k) out[g0 + k] = NAN;
            return ORACLE_OK;
        }
        delta = W - n;
        len = W;
    }
    const int64_t front = delta / 2;
    for (int64_t k = 0; k < len; ++k) prob[k] = 0.0;
    for (int64_t i = 0; i + W <= len; i += step) {
        for (int32_t k = 0; k < W; ++k) {
            const int64_t pos = i + k - front;
            items[k] = (pos >= 0 && pos < n) ? g0 + pos : -1;
        }
        score_items(m, ctx, gene_ptr, attr_idx, items, W);
        forward_backward(m, ctx, W);
        for (int32_t k = 0; k < W; ++k) {
            const double p = marginal(ctx, m->L, k, pos_label);
            /* numpy.maximum(probabilities[win], marginals): NaN propagates */
            if (p > prob[i + k] || p != p) prob[i + k] = p;
        }
        ++*windows_done;
    }
    for (int64_t k = 0; k < n; ++k) out[g0 + k] = prob[front + k];
    return ORACLE_OK;
}

typedef struct {
    const oracle_model *m;
    const int64_t *contig_ptr, *gene_ptr;
    const int32_t *attr_idx;
    int64_t C;
    int32_t W, step, pad, pos_label;
    double *out;
    int64_t next;          /* shared work cursor (contig index), guarded by lock */
    int64_t windows;       /* total windows processed */
    pthread_mutex_t lock;
    int rc;
} job_t;

#define JOB_CHUNK 16

static void *worker(void *arg) {
    job_t *job = (job_t *)arg;
    oracle_ctx ctx; memset(&ctx, 0, sizeof(ctx));
    int64_t *items = (int64_t *)malloc(sizeof(int64_t) * (size_t)job->W);
    double *prob = NULL; int64_t prob_cap = 0; int64_t windows = 0;
    int rc = items ? ctx_reserve(&ctx, job->W, job->m->L) : ORACLE_ENOMEM;
    while (rc == ORACLE_OK) {
        pthread_mutex_lock(&job->lock);
        const int64_t c0 = job->next;
        job->next = c0 + JOB_CHUNK;
        pthread_mutex_unlock(&job->lock);
        if (c0 >= job->C) break;
        const int64_t c1 = (c0 + JOB_CHUNK < job->C) ? c0 + JOB_CHUNK : job->C;
        for (int64_t c = c0; c < c1 && rc == ORACLE_OK; ++c) {
            const int64_t g0 = job->contig_ptr[c], n = job->contig_ptr[c + 1] - g0;
            if (n <= 0) continue;
            const int64_t need = (n > job->W) ? n : job->W;
            if (need > prob_cap) {
                free(prob);
                prob = (double *)malloc(sizeof(double) * (size_t)need);
                prob_cap = prob ? need : 0;
                if (!prob) { rc = ORACLE_ENOMEM; break; }
            }
            rc = contig_windowed(job->m, &ctx, items, prob, job->gene_ptr, job->attr_idx, g0, n, job->W,
                                 job->step, job->pad, job->pos_label, job->out, &windows);
        }
    }
    pthread_mutex_lock(&job->lock);
    job->windows += windows;
    if (rc != ORACLE_OK) job->rc = rc;
    pthread_mutex_unlock(&job->lock);
    free(items); free(prob); ctx_release(&ctx);
    return NULL;
}

/*
 * Whole batch: per-gene P(label pos_label) for C contigs.  contig_ptr[C+1] indexes genes,
 * gene_ptr[G+1] indexes attr_idx.  nthreads <= 1 runs on the calling thread; otherwise contigs are
 * handed out in chunks to nthreads pthreads (the reference itself is single-threaded; threads over
 * contigs are how a user would scale it on the host).  windows_out (optional) receives the number
 * of windows evaluated (the `total` of gecco/crf/__init__.py:239 when step == 1).
 */
int oracle_marginals_windowed(const double *state_w, int32_t A, int32_t L, const double *trans_w,
                              int32_t pos_label, const int64_t *contig_ptr, const int64_t *gene_ptr,
                              const int32_t *attr_idx, int64_t C, int32_t window, int32_t step,
                              int32_t pad, int32_t nthreads, double *out, int64_t *windows_out) {
    if (window <= 0 || step <= 0 || step > window) return ORACLE_EINVAL; /* _meta.py:127-130 */
    if (pos_label < 0 || pos_label >= L || C < 0 || !out) return ORACLE_EINVAL;
    oracle_model m;
    int rc = model_init(&m, state_w, A, L, trans_w);
    if (rc) return rc;
    job_t job;
    memset(&job, 0, sizeof(job));
    job.m = &m; job.contig_ptr = contig_ptr; job.gene_ptr = gene_ptr; job.attr_idx = attr_idx;
    job.C = C; job.W = window; job.step = step; job.pad = pad; job.pos_label = pos_label; job.out = out;
    pthread_mutex_init(&job.lock, NULL);
    if (nthreads <= 1) {
        worker(&job);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
        if (!th) { free(m.exp_trans); return ORACLE_ENOMEM; }
        int started = 0;
        for (; started < nthreads; ++started)
            if (pthread_create(&th[started], NULL, worker, &job) != 0) break;
        if (started == 0) worker(&job);
        for (int k = 0; k < started; ++k) pthread_join(th[k], NULL);
        free(th);
    }
    pthread_mutex_destroy(&job.lock);
    if (windows_out) *windows_out = job.windows;
    free(m.exp_trans);
    return job.rc;
}

int oracle_version(void) { return 1; }
