/*
 * ddb200.h -- C ABI of the B200-native generate -> solve -> label hot path of rodrgo/deep_dantzig.
 *
 * The reference has no FFI; its boundary for this path is a Python class surface.
 * Each entry point below replaces the per-instance Python/Gurobi code cited beside it with one batched call.
 * Plain pointers and sizes only.  `*_dev` entry points take DEVICE pointers and a CUDA stream handle
 * (cudaStream_t passed as void*, NULL = default stream) and are asynchronous on that stream; `*_host`
 * entry points take HOST pointers, do the host<->device copies themselves and return when results are in
 * the caller's buffers.
 *
 * Every function returns 0 on success and a negative DDB_E* code on failure; ddb_last_error() gives the
 * message for the calling thread.  Per-instance solver outcomes are reported in `status[]` using the Gurobi
 * code table the reference switches on (src/data/gurobi_lp.py:447-461).
 */
#ifndef DDB200_H
#define DDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDB_ABI_VERSION 2

/* return codes */
#define DDB_OK            0
#define DDB_EINVAL       -1   /* bad argument (NULL pointer, m < 1, ...)                      */
#define DDB_ECUDA        -2   /* CUDA runtime / launch error, see ddb_last_error()            */
#define DDB_EUNSUPPORTED -3   /* shape not supported by any kernel (see ddb_last_error())     */
#define DDB_ENOMEM       -4

/* per-instance status codes == Gurobi's (src/data/gurobi_lp.py:447-461) */
#define DDB_ST_LOADED           1
#define DDB_ST_OPTIMAL          2
#define DDB_ST_INFEASIBLE       3
#define DDB_ST_INF_OR_UNBD      4
#define DDB_ST_UNBOUNDED        5
#define DDB_ST_ITERATION_LIMIT  7
#define DDB_ST_NUMERIC         12

/* The reference's active-constraint threshold (src/data/gurobi_lp.py:437). */
#define DDB_DEFAULT_THRESHOLD 1e-7

typedef struct ddb_ctx ddb_ctx;

int         ddb_abi_version(void);
const char *ddb_last_error(void);

/* One context per (process, device): owns the work-queue counters, kernel scratch (parked rows, crash inverses, fused-mode
 * instance slabs -- all L2-resident) and the pinned staging ring of the *_host entry points.  Entry points that touch this
 * state take the context's mutex, so a context may be shared between threads and streams: launches that use the scratch
 * are ordered after one another by an event, whatever stream they were issued on. */
int ddb_create(int device, ddb_ctx **out);
int ddb_destroy(ddb_ctx *ctx);
/* sm_count, compute capability, opt-in shared memory per block (bytes) of the context's device. */
int ddb_device_info(ddb_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, int64_t *smem_optin);

/* Which kernel family ddb_solve_label_* will use for an (m, n) shape: 0 = tableau in the register file
 * (row-per-thread kernel, falling back to the other register kernels for shapes it does not cover),
 * 1 = shared-memory tableau, 2 = global-memory (L2/HBM streamed) tableau, 6 = thread-block-cluster kernel (the live
 * tableau in the distributed shared memory of 1-8 SMs: shapes beyond one SM such as (500,250)), 7 = tableau in the
 * register file, rows over lanes and columns over warps (72 <= n <= 100, m - n <= 128, e.g. (200,100), with or without a
 * row mask, with or without the in-solver generator); <0 = error. */
int ddb_solve_plan(ddb_ctx *ctx, int m, int n);
/* Force a kernel family for testing (-1 = automatic).  Beyond 0..2: 4 = warp-tiled register kernel (the fallback of plan 0
 * for 100 < n <= 111).  3 (2-D register tile) and 5 (software-pipelined rows) are measured negative results that only the
 * `make experiments` build of the library contains; the shipped library answers DDB_EUNSUPPORTED for them. */
int ddb_set_solve_plan(ddb_ctx *ctx, int plan);

/*
 * (1) GENERATE -- replaces the serial loop RandomLPDataset._generate_problems + the generator part of
 * create_lp_problem (src/data/randomlp_dataset.py:58-63, 76-86) in throughput mode.
 * Instance i = first_instance + k (k in [0,B)) is a pure function of (key, i): Philox4x32-10 with
 * key = (key_lo, key_hi), counter = (element, stream, i_lo, i_hi); Box-Muller normals.
 *   A[k,:,:] ~ N(0,1) (row-major m x n, each entry kept with probability `density`, else 0 -- density 1.0
 *   is the reference's distribution), x0 ~ N(0,I_n), b = A x0 + |N(0,I_m)|, c = |N(0,I_n)|.
 * x0 may be NULL.
 */
int ddb_generate_dev(ddb_ctx *ctx, uint64_t key, int64_t first_instance, int64_t B, int m, int n,
                     double density, double *A, double *b, double *c, double *x0, void *stream);

/*
 * (2) SOLVE + LABEL -- replaces LinProg(A,b,c,'min',['<']*m) + optimize() + get_statuscode() +
 * get_active_constraints() (src/data/gurobi_lp.py:11-29, 428-465) and the label assembly of
 * create_lp_problem (src/data/randomlp_dataset.py:88-106) for a batch of B LPs
 *     min c'x  s.t.  A x <= b,  x free.
 * Inputs  : A[B,m,n], b[B,m], c[B,n] fp64 row-major; threshold (the reference uses 1e-7);
 *           row_mask[B,m] (nullable): rows with mask 0 are left out of the LP that is solved (the
 *           "reduced LP" of BASELINE.json config 4); labels / violations are still evaluated on all m rows.
 * Outputs : status[B]; x[B,n], obj[B] (defined when status == 2); labels[B,m] = 1 where
 *           |b - A x| <= threshold, all 0 when status != 2 (randomlp_dataset.py:96-102);
 *           n_active[B]; pivots[B,4] = {crash, phase-1, phase-2, total} pivot counts;
 *           ties[B] = rows whose |slack| lies in [threshold/10, threshold*10] plus rows whose label disagrees
 *           with final-basis membership; violations[B] (nullable) = rows with slack < -threshold at the
 *           returned x (non-zero only for reduced LPs: a certificate that the reduced optimum is the full one).
 * Any output pointer except status and labels may be NULL.
 */
int ddb_solve_label_dev(ddb_ctx *ctx, int64_t B, int m, int n,
                        const double *A, const double *b, const double *c,
                        double threshold, const uint8_t *row_mask,
                        int32_t *status, double *x, double *obj, uint8_t *labels,
                        int32_t *n_active, int32_t *pivots, int32_t *ties, int32_t *violations,
                        void *stream);

/* Same contract, HOST buffers; H2D / D2H copies are done inside, in chunks of ~192 MB with three chunks in flight.
 * Page-locked caller memory (cudaHostAlloc / cudaHostRegister) is DMA'd in place.  PAGEABLE caller memory (a plain numpy
 * array, malloc) is first copied by the host's cores (DDB_COPY_THREADS, default min(cores, 16)) into the context's pinned staging
 * ring and DMA'd from there, and results come back through pinned staging the same way -- so the copies of one chunk
 * overlap the solve of another for any caller, not only for one that pinned its buffers. */
int ddb_solve_label_host(ddb_ctx *ctx, int64_t B, int m, int n,
                         const double *A, const double *b, const double *c,
                         double threshold, const uint8_t *row_mask,
                         int32_t *status, double *x, double *obj, uint8_t *labels,
                         int32_t *n_active, int32_t *pivots, int32_t *ties, int32_t *violations);

/*
 * (3) FUSED GENERATE -> SOLVE -> LABEL -- the whole loop RandomLPDataset._generate_problems -> create_lp_problem
 * (src/data/randomlp_dataset.py:58-63, 65-128) for instances first_instance..+B of stream `key`; the caller never
 * materialises A, b, c.  Same outputs as (2).  Two implementations, same bits and same results:
 *   mode 2 (automatic choice, the measured-faster one): generator kernel (1) into context scratch, chunk by chunk, chunk
 *          i + 1 generated on a side stream while chunk i is solved;
 *   mode 1: ONE kernel launch -- the thread block that solves instance i draws it first into a per-block slab that stays
 *          in L2, so A never travels through HBM (even n; shapes of the register-resident kernels, plans 0 and 7).
 * ddb_set_fused_mode(ctx, 0 | 1 | 2) selects (0 = automatic).  A_out/b_out/c_out (device; all three or none) receive the
 * instances when the caller wants them.
 */
int ddb_set_fused_mode(ddb_ctx *ctx, int mode);
int ddb_generate_solve_label_dev(ddb_ctx *ctx, uint64_t key, int64_t first_instance, int64_t B, int m, int n,
                                 double density, double threshold,
                                 int32_t *status, double *x, double *obj, uint8_t *labels,
                                 int32_t *n_active, int32_t *pivots, int32_t *ties, int32_t *violations,
                                 double *A_out, double *b_out, double *c_out, void *stream);
/* (3) with HOST output arrays (what the reference's loop hands its caller: labels, status, objective, x ...): nothing
 * travels host -> device, results come back chunk by chunk while the next chunk is being solved. */
int ddb_generate_solve_label_host(ddb_ctx *ctx, uint64_t key, int64_t first_instance, int64_t B, int m, int n,
                                  double density, double threshold,
                                  int32_t *status, double *x, double *obj, uint8_t *labels,
                                  int32_t *n_active, int32_t *pivots, int32_t *ties, int32_t *violations);

/*
 * (4) CLASSIFIER FORWARD -- replaces Model.forward (src/ml/models/s2v.py:45-54; _forward_complete :124-187,
 * _forward_bipartite :253-323) and the per-instance batch loop of ml/utils.py:3-25 for B instances of one shape.
 * graph: 0 = 'complete', 1 = 'bipartite'.  p = embedding dimension, T = rounds_s2v.  A, b, c are the fp64 arrays of
 * (2); in_loss is every row (the random-LP feature adapter).
 * params: flat fp32 block in the reference's state_dict order and shapes (s2v.py:60-89 / 189-216):
 *   complete : t0[p] t1[p] t2rr t2rc t2cr t3rr t3rc t3cr [p*p each] t4rr[p] t4rc[p] t4cr[p] t6r t6c t7 [p*p] t8[2*2p]
 *   bipartite: t0[p] t1c[p*4] t1v[p] t2c t2v t3c t3v [p*p] t4c[p] t4v[p] t6c t6v t7 [p*p] t8[2*(2p+4)]
 * Outputs: logp[B,m,2] = log_softmax(scores) (what forward returns), probs[B,m,2] (Model.probs; nullable).
 */
int ddb_s2v_param_count(int graph, int p);
int ddb_s2v_forward_dev(ddb_ctx *ctx, int graph, int64_t B, int m, int n, int p, int T,
                        const double *A, const double *b, const double *c, const float *params,
                        float *logp, float *probs, void *stream);
/* The same forward for items that are not plain random LPs -- MPS / PLNN items (src/data/gurobi_lp.py:127-187, 326-366)
 * with equality rows and bound rows: row_ineq[B,m], row_bound[B,m] (0 / 1) are the per-row node flags the reference keeps
 * in c_feats[:, 0] (is_inequality) and c_feats[:, 2] (is_bound) of a bipartite item, row_ineq alone the node_features of
 * the row nodes of a complete item (row_bound is ignored for graph 0).  NULL flags = the random-LP values (1 and 0), i.e.
 * ddb_s2v_forward_dev.  Bipartite items carry both flags or none; they run on the general-adjacency kernel. */
int ddb_s2v_forward_flags_dev(ddb_ctx *ctx, int graph, int64_t B, int m, int n, int p, int T,
                              const double *A, const double *b, const double *c, const float *params,
                              const uint8_t *row_ineq, const uint8_t *row_bound,
                              float *logp, float *probs, void *stream);

/*
 * (5) CLASSIFIER LOSS + GRADIENT -- replaces the per-instance accumulation loop of train_net
 * (src/ml/train.py:59-66: for every instance  loss = criterion(model(x), y); loss.backward()) with its criterion
 * NLLLoss(weight=[w0, w1], size_average=False) (src/benchmark.py:70-75) for a batch of B instances of one shape.
 * labels[B,m] u8 (0/1) are the active-constraint labels of (2).  Outputs (device, overwritten):
 *   grad[ddb_s2v_param_count(graph, p)] = d loss / d params in the flat order of (4), summed over the batch;
 *   loss = summed weighted negative log-likelihood (fp64 accumulation of fp32 terms);
 *   not_dense = 1 if an instance has a zero in A and graph == 1 (informational: the streaming kernel covers the reference's
 *   dense random LPs and flags such instances, a general-adjacency kernel then adds exactly those).  Always 0 for graph 0.
 * graph 1 ('bipartite', the reference's default, benchmark.py:166): p <= 64.
 * graph 0 ('complete'): tcgen05 Gram kernel (relu row sums of W = G G^T) + a forward/backward kernel on them; needs
 *   m + 1 <= 256, p <= 64 and T p m floats of embeddings in shared memory, else DDB_EUNSUPPORTED ("do not fit").
 */
int ddb_s2v_loss_grad_dev(ddb_ctx *ctx, int graph, int64_t B, int m, int n, int p, int T,
                          const double *A, const double *b, const double *c, const float *params,
                          const uint8_t *labels, float w0, float w1,
                          float *grad, double *loss, int32_t *not_dense, void *stream);
/* The same step for MPS / PLNN items: per-row node flags as in ddb_s2v_forward_flags_dev, and a label value of 2 marks a row
 * outside the item's in_loss set (it contributes neither loss nor gradient).  Bipartite items run on the general-adjacency
 * kernel. */
int ddb_s2v_loss_grad_flags_dev(ddb_ctx *ctx, int graph, int64_t B, int m, int n, int p, int T,
                                const double *A, const double *b, const double *c, const float *params,
                                const uint8_t *labels, const uint8_t *row_ineq, const uint8_t *row_bound,
                                float w0, float w1, float *grad, double *loss, int32_t *not_dense, void *stream);

/*
 * (6) CLASSIFIER EVALUATION METRICS -- replaces the host side of the per-epoch evaluation: the recall-1 threshold
 * taken from sklearn's roc_curve (src/ml/train.py:118-150; the first ROC threshold with TPR == 1.0 is the smallest
 * predicted probability of a positive) and the confusion-matrix loop of performance() (train.py:174-246) /
 * get_accuracy (src/ml/test.py:10-54).  One streaming pass over N = B*m constraint nodes.
 * Inputs : probs[N,2] (Model.probs), logp[N,2] (nullable: loss is 0 without it), labels[N] u8, thresh, class weights.
 * Output : out[8] fp64 on the device = { tp, fp, tn, fn  (predicted positive <=> probs[:,1] >= thresh),
 *          min over positives of probs[:,1] (+inf if none), weighted NLL sum, #positives, #negatives }.
 */
int ddb_s2v_metrics_dev(ddb_ctx *ctx, int64_t N, const float *logp, const float *probs, const uint8_t *labels,
                        float thresh, float w0, float w1, double *out, void *stream);

/* Number of kernels this library has launched on the context since creation (bench.py's gpu_launches). */
int64_t ddb_launch_count(ddb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* DDB200_H */
