/* dmpc.h — C-ABI of the B200-native Felsenstein-pruning likelihood engine (libdmpc.so).
 *
 * Drop-in boundary for the ONE hot path of villandre/DMphyClus: the likelihood that
 * `DMphyClusChain` / `logLikFromClusInd` evaluate over the AugTree.  Each entry point replaces one
 * Rcpp-exported function of the reference (`src/clusterFunArmadillo.cpp`, registered in
 * `src/RcppExports.cpp:179-192`, called from `R/RcppExports.R:4-46`); the citation beside each
 * declaration is the reference interface it stands in for.  INTEGRATION.md shows the Rcpp shim a
 * maintainer adds so that the R API stays unchanged.
 *
 * Conventions (all taken from what R hands the reference):
 *   - node ids, matrix-list indices and edge matrices are R 1-BASED;
 *   - an edge matrix is int32 [2*nEdgeRows], COLUMN-major (parents first, then children), nEdgeRows = 2N-2;
 *   - a transition-matrix list is double [R*16]: R column-major 4x4 matrices, P(i,j) = Pr(parent i -> child j),
 *     used as P * v (IntermediateNode.cpp:57);
 *   - limProbs is double [4] in the same state order as the matrices (names(limProbs));
 *   - the alignment is the packed form of getConvertedAlignment's result: uint8 [numTips*numLoci], row-major
 *     by tip, bit j of a cell set when state j is compatible with the observed symbol.
 * Everything is plain pointers and sizes; no C++/torch types cross the boundary.  All calls are synchronous:
 * on return the log-likelihood has been computed on the GPU and copied back.  There is no CPU fallback: every
 * entry point fails with DMPC_ERR_CUDA if no sm_100-class device is usable.
 *
 * Errors: functions return DMPC_OK (0) or a negative status; dmpc_last_error() gives the message for the calling
 * thread.  The Rcpp shim turns a non-zero status into Rcpp::stop() (the reference throws Rcpp::exception,
 * clusterFunArmadillo.cpp:136).
 */
#ifndef DMPC_H
#define DMPC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dmpc_tree dmpc_tree; /* opaque; replaces XPtr<AugTree> (clusterFunArmadillo.cpp:57) */

enum {
  DMPC_OK = 0,
  DMPC_ERR_ARG = -1,     /* malformed input (edge matrix, sizes, null pointer, unknown symbol) */
  DMPC_ERR_CUDA = -2,    /* CUDA runtime/driver failure, or no usable device */
  DMPC_ERR_NO_MOVE = -3, /* no NNI candidate exists (the reference aborts inside GSL here) */
  DMPC_ERR_STATE = -4    /* call sequence error (e.g. use after dmpc_final_deallocate) */
};

/* Options fixed at handle creation (the R signature has no room for them; the shim reads env vars). */
typedef struct dmpc_options {
  int device;        /* CUDA device ordinal; -1 = current device */
  int quirk_carry;   /* 1 (default): reproduce AugTree.cpp:306 — the fp32 exponent vector is NOT reset between
                        loci, so sites are coupled through a sequential fp32 chain.  0: textbook per-site reduction */
  int site_begin;    /* this handle evaluates global sites [site_begin, site_begin + numLoci) of an alignment of */
  int sites_total;   /* sites_total sites (multi-GPU site sharding); single GPU: 0 and numLoci */
  int fuse_tips;     /* small-clade fusion: vertices with at most this many tips below them are never written to
                        HBM but recomputed from the tip masks by their nearest stored ancestor (default and max 15;
                        <= 1 stores every vertex).  Results do not depend on it */
  int fuse_exact;    /* 1: fused clades keep full per-vertex rescale bookkeeping from the start (the engine switches
                        to it by itself the first time a fused clade could have rescaled) */
  void* comm;        /* dmpc_comm* of a site-sharded run (or NULL): deferred evaluations then exchange their chunk sums
                        through NVLink peer mailboxes inside the reduction kernel */
  int deferred;      /* 1: start in deferred-reduction mode (see dmpc_set_deferred): dmpc_loglik_create prunes and
                        gathers but leaves the root reduction to dmpc_shard_chain / dmpc_shard_finish */
  int chain_cert;    /* 1 (default): where it can be PROVEN that the reference's sequential fp32 exponent chain is in its
                        one-dominant-category regime, the per-site terms are computed in parallel from the closed form
                        (bit-identical); 0: always walk the chain literally (tests compare the two) */
  int n_devices;     /* > 1: single-process multi-GPU — the handle shards its sites over devices[0 .. n_devices) */
  int devices[8];    /*      and every entry point fans out to the shards (the R shim reads DMPC_DEVICES) */
} dmpc_options;
void dmpc_default_options(dmpc_options* o);

const char* dmpc_last_error(void);
const char* dmpc_version(void);

/* getConvertedAlignment (clusterFunArmadillo.cpp:64-113): chars [numTips*numLoci] row-major, lower-case
 * DNA/IUPAC; states = the 4 names(limProbs).  Returns the number of cells holding an unknown symbol
 * (their mask is 0; dmpc_loglik_create rejects such alignments). */
long dmpc_convert_alignment(const char* chars, long nCells, const char* states, uint8_t* maskOut);

/* logLikCpp (clusterFunArmadillo.cpp:36-62): builds the tree, uploads the alignment, evaluates. */
int dmpc_loglik_create(const int32_t* edge, int nEdgeRows, const double* clusterMRCAs, int nClusterMRCAs,
                       const double* limProbs, const double* withinTransMats, const double* betweenTransMats,
                       int numRateCats, const uint8_t* alignmentBin, int numTips, int numLoci,
                       int withinMatListIndex, int betweenMatListIndex, const dmpc_options* opts,
                       dmpc_tree** treeOut, double* logLikOut);

/* newBetweenTransProbsLogLik (clusterFunArmadillo.cpp:117-138) */
int dmpc_new_between_transprobs_loglik(dmpc_tree* t, const double* withinTransMats, const double* newBetweenTransMats,
                                       int newBetweenMatListIndex, const double* limProbs, double* logLikOut);
/* newWithinTransProbsLogLik (clusterFunArmadillo.cpp:142-164) */
int dmpc_new_within_transprobs_loglik(dmpc_tree* t, const double* newWithinTransMats, const double* betweenTransMats,
                                      int newWithinMatListIndex, const double* limProbs, double* logLikOut);
/* withinClusNNIlogLik (clusterFunArmadillo.cpp:168-203); edgeOut: int32 [2*(2N-2)] column-major, pre-order */
int dmpc_within_clus_nni_loglik(dmpc_tree* t, const double* withinTransMats, const double* betweenTransMats,
                                int mrcaOfClusForNNI, int numMovesNNI, const double* limProbs, double* logLikOut,
                                int32_t* edgeOut);
/* betweenClusNNIlogLik (clusterFunArmadillo.cpp:207-242) */
int dmpc_between_clus_nni_loglik(dmpc_tree* t, const double* withinTransMats, const double* betweenTransMats,
                                 const double* clusterMRCAs, int nClusterMRCAs, int numMovesNNI,
                                 const double* limProbs, double* logLikOut, int32_t* edgeOut);
/* clusSplitMergeLogLik (clusterFunArmadillo.cpp:246-276): n == 1 splits that cluster MRCA, n > 1 merges */
int dmpc_clus_split_merge_loglik(dmpc_tree* t, const double* withinTransMats, const double* betweenTransMats,
                                 const int32_t* clusMRCAsToSplitOrMerge, int n, const double* limProbs,
                                 double* logLikOut);
/* RestorePreviousConfig (clusterFunArmadillo.cpp:280-284; AugTree.cpp:348-366).  edge is read only when nniMove. */
int dmpc_restore_previous_config(dmpc_tree* t, const int32_t* edge, int nEdgeRows, int withinMatListIndex,
                                 int betweenMatListIndex, int nniMove, const int32_t* clusterMRCAs,
                                 int nClusterMRCAs, int splitMergeMove);
/* finalDeallocate (clusterFunArmadillo.cpp:288-294) — also frees the tree itself, which the reference leaks */
int dmpc_final_deallocate(dmpc_tree* t);
/* getMaxMapSizeEstimate (clusterFunArmadillo.cpp:298-302): same arithmetic, kept for signature parity */
double dmpc_get_max_map_size_estimate(double allowedSizeInMegs, unsigned numRateCats);
/* checkAndCullMap (clusterFunArmadillo.cpp:305-332): the dense HBM layout has no dictionary to cull; no-op */
int dmpc_check_and_cull_map(dmpc_tree* t, unsigned allowedNumberOfElements, float cullProportion,
                            const double* withinTransMats, const double* betweenTransMats, const double* limProbs);

/* ---------------------------------------------------------------- beyond the reference's 11 entry points */

/* Re-evaluate everything with the current topology/flags/matrices (InvalidateAll + ComputeLoglik); the bench's
 * device-resident "step".  Does not touch the rollback state of the last proposal. */
int dmpc_recompute_all(dmpc_tree* t, const double* withinTransMats, const double* betweenTransMats,
                       const double* limProbs, double* logLikOut);

/* A batch of cluster proposals evaluated against the COMMITTED state, which is left untouched: logLikOut[i] is what
 * dmpc_clus_split_merge_loglik would return for proposal i (followed by a reject).  Proposals whose paths involve no
 * rescaling are evaluated together — one launch walks every proposal's path to the root with the path partial in
 * registers, reading only the off-path siblings; the rest goes through the ordinary path one by one.  *nBatchedOut =
 * how many took the batched kernel.  Matrices must be the committed ones.  Call it only when no proposal is waiting
 * for its dmpc_restore_previous_config: the one-by-one path reuses the rollback slots (a proposal made before the call
 * can no longer be restored afterwards), and a fused clade that turns out to rescale switches the handle to exact
 * mode for good (dmpc_stats.exact_mode), exactly as an ordinary proposal would. */
enum { DMPC_PROPOSAL_SPLIT = 0, DMPC_PROPOSAL_MERGE = 1 };
typedef struct dmpc_proposal {
  int32_t kind;   /* DMPC_PROPOSAL_SPLIT: a = cluster MRCA to split; DMPC_PROPOSAL_MERGE: a, b = sibling cluster MRCAs */
  int32_t a, b;   /* 1-based vertex ids */
} dmpc_proposal;
int dmpc_eval_proposals(dmpc_tree* t, const double* withinTransMats, const double* betweenTransMats,
                        const double* limProbs, const dmpc_proposal* props, int nProps, double* logLikOut,
                        int* nBatchedOut);

/* Counters since creation: node-site-rate partial updates performed, kernels launched, seconds of device time of
 * the last evaluation (CUDA events on the engine's stream). */
typedef struct dmpc_stats {
  unsigned long long updates;
  unsigned long long kernel_launches;
  unsigned long long evaluations;
  double last_eval_ms;      /* whole evaluation: prune levels + root reduction */
  double last_prune_ms;     /* the prune_level launches only */
  unsigned long long last_updates;   /* updates of the last evaluation */
  unsigned long long last_prune_launches;
  int last_levels;
  int exponent_capacity;    /* rows of the per-site rescale list (grows on demand) */
  unsigned long long device_bytes;   /* bytes of HBM held by this handle */
  int last_active_tiles;    /* 128-site tiles that held a rescaled site in the last evaluation (0: the fp32
                               exponent chain is the identity on this handle's sites) */
  int num_tips, num_loci, num_rate_cats;   /* shape of the handle (the shim sizes edge outputs from it) */
  unsigned long long last_algorithmic_bytes;   /* compulsory HBM bytes of the last evaluation's pruning: stored
                               partials written + stored operands read (32 B each) + 1 B per tip operand */
  int last_stored_vertices, last_fused_vertices;   /* vertices written to HBM / recomputed in registers */
  int exact_mode;           /* 1 once fused clades run with full rescale bookkeeping */
  double last_host_plan_us;      /* host time of the last evaluation: building the plan (schedule, tasks, tokens) */
  double last_host_submit_us;    /* ... uploading it and launching every kernel */
  double last_host_total_us;     /* ... whole evaluate(), including the wait for the result */
  double create_setup_us;        /* dmpc_loglik_create: host time before the alignment upload is issued (tree, buffers) */
  double create_upload_ms;       /* ... device time of the alignment upload + re-pitch (CUDA events) */
  int last_chain_sites, last_chain_rows;   /* rescaled sites / exponent rows the sequential chain kernel walked */
  int chain_group;          /* rows per group the chain walker will use next (4, or 1 for sparse exponent lists) */
  unsigned long long plan_cache_hits;   /* full evaluations that replayed the device-resident plan */
  unsigned long long graph_replays;     /* ... of which submitted as one CUDA graph launch */
  int last_walk_steps;      /* vertices of the last evaluation's trailing chain, evaluated by the walk kernel with the
                               running partial in registers (0: no chain of at least 4 single-task levels) */
  int last_materialised;    /* fused-size clades hanging off that chain, evaluated once as stored tasks beforehand */
  int last_cert;            /* how the cross-locus exponent chain of the last evaluation was resolved: 0 = no site
                               rescaled, 1 = by certificate (last_chain_sites = its literal prefix), 2 = walked literally */
  int last_cert_rate;       /* the dominant rate category of the certified regime (-1: none) */
  double prune_ms_total;    /* device time of the pruning launches of every evaluation so far (CUDA events on the engine's
                               stream around the round launches; read without a synchronisation once the result is in) */
  unsigned long long last_fp64_ops;   /* DMUL + DADD instructions (per lane) the last evaluation's pruning had to issue: a 4x4
                               mat-vec = 16 + 12, the product of the two child terms = 4, tip terms come from a table */
} dmpc_stats;
int dmpc_get_stats(dmpc_tree* t, dmpc_stats* out);
/* Latency breakdown of the last evaluation's root stage from %globaltimer stamps the root kernels leave in the status
 * block (microseconds; -1 where a phase did not run): usOut[0] gather, first block to last block; [1] the gather's
 * cross-rank exchange (wait for the peers' messages) or the quiet reduction; [2] chain certificate + literal prefix;
 * [3] gap until the final kernel's first block; [4] final kernel, per-site terms; [5] its exchange / reduction.
 * (Multi-device handle: shard 0's.) */
int dmpc_get_root_timeline(dmpc_tree* t, double* usOut);
/* Device-side timing of a region of calls: CUDA events recorded on the engine's own stream (which
 * torch.cuda.Event cannot see).  end returns the elapsed milliseconds between the two records. */
int dmpc_timer_begin(dmpc_tree* t);
int dmpc_timer_end(dmpc_tree* t, double* msOut);

/* (The host-only scheduling hooks the CPU test-suite drives — dmpc_plan_selfcheck, dmpc_hp_* — are declared in
 * include/dmpc_test.h: they are exported by the same library but are not part of the product interface.) */

/* Introspection for parity tests: the committed partial of a vertex (0-based id >= numTips) as
 * sol [numLoci][R][4] and expo [numLoci][R] (the oracle's layout), and the topology/flags. */
int dmpc_get_partial(dmpc_tree* t, int vertex, double* sol, float* expo);
int dmpc_get_topology(dmpc_tree* t, int32_t* parent, int32_t* child0, int32_t* child1, int32_t* within);
/* per-site terms of the last evaluation: rateAveragedLogLiks (AugTree.cpp:322), [numLoci] */
int dmpc_get_site_logliks(dmpc_tree* t, double* out);

/* Multi-GPU site sharding (one handle per rank, contiguous site blocks).  With quirk_carry and R > 1 the
 * fp32 exponent chain crosses shard boundaries: a rank's chain starts from the previous rank's carry.
 *   (an entry point called on a deferred handle — dmpc_options.deferred / dmpc_set_deferred — prunes the dirty
 *   vertices and gathers the per-site quantities, and stops there unless peer mailboxes are attached;)
 *   dmpc_shard_chain:  runs this shard's chain from carryIn[R] (fp32) and returns carryOut[R];
 *   dmpc_shard_finish: returns the deterministic partial sums of this shard's sites in fixed 64-site chunks
 *                      (chunkSums[nChunks]); the caller all-gathers them and adds them in chunk order, so the
 *                      total is bit-identical for any number of ranks.
 * With deferral on, the entry points above return logLik = NaN and leave the reduction to these calls. */
int dmpc_set_deferred(dmpc_tree* t, int on);
int dmpc_shard_chain(dmpc_tree* t, const float* carryIn, float* carryOut);
int dmpc_shard_active(dmpc_tree* t);     /* 1 when a site of this handle rescaled in the last evaluation */
int dmpc_shard_num_chunks(dmpc_tree* t);
int dmpc_shard_finish(dmpc_tree* t, double* chunkSums);
/* Peer mailboxes for the fused exchange (one process per GPU of ONE box, peer access over NVLink/NVSwitch).
 *   dmpc_comm_create:  allocates this rank's mailbox (room for maxChunksPerRank chunk sums per sender) and returns the
 *                      64-byte CUDA IPC handle other ranks open it with;
 *   dmpc_comm_connect: opens the world mailboxes from their handles (allHandles = world x 64 bytes, in rank order);
 * then pass the comm in dmpc_options.comm (with deferred = 1).  After each deferred evaluation
 *   dmpc_shard_fast_result: *ok = 1 and *logLik set when the in-kernel exchange completed it (no rank held a rescaled
 *                      site); *ok = 0: run the host protocol (dmpc_shard_chain / dmpc_shard_finish + a collective). */
typedef struct dmpc_comm dmpc_comm;
int dmpc_comm_create(int device, int rank, int world, int maxChunksPerRank, dmpc_comm** out, void* ipcHandle64);
int dmpc_comm_connect(dmpc_comm* c, const void* allHandles);
int dmpc_comm_destroy(dmpc_comm* c);
int dmpc_shard_fast_result(dmpc_tree* t, int* ok, double* logLik);

/* The contiguous block of sites [*begin, *end) that rank / shard `rank` of `world` owns in an alignment of sitesTotal sites:
 * equal blocks, rounded up to the 64-site reduction chunk (so that chunk boundaries — and the summation tree — do not depend
 * on the number of ranks); trailing ranks of a short alignment get an empty block.  This is the layout of a multi-device
 * handle (dmpc_options.n_devices) and the one dmphyclus_b200/sharded.py uses for process-per-GPU runs. */
void dmpc_shard_range(int sitesTotal, int rank, int world, int* begin, int* end);

/* Host-side twin of the device's final reduction: adds n chunk sums in the same fixed order as the single-GPU
 * kernel (256 strided accumulators, then a binary tree), so 1-GPU and N-GPU totals agree bit for bit. */
double dmpc_reduce_chunks(const double* chunkSums, int n);

#ifdef __cplusplus
}
#endif
#endif /* DMPC_H */
