// ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
//
// C-ABI driver around the reference's OWN, UNMODIFIED translation units
// (/root/reference/src/{clusterFunArmadillo,AugTree,IntermediateNode,InputNode}.cpp), compiled from
// where they lie against the header shim in oracle/shim (no R, Rcpp, Armadillo, GSL or Boost in this
// image).  Output: oracle/_ref/libref.so (git-ignored).  It builds the shim's "R objects" from plain
// arrays, calls the reference's Rcpp-exported functions (clusterFunArmadillo.cpp:36-332) and reads
// the results back.  It is the second, independent oracle used to pin oracle/dmpc_oracle.cpp
// (tests/test_oracle_vs_ref.py) and the `cpu_baseline.kind == "reference"` arm of bench.py.
// Built twice (oracle/Makefile):
//   _ref/libref.so       with the reference's own translation units (default);
//   _ref/libshimtest.so  -DHARNESS_DMPC: with rpkg/src/clusterFunArmadillo.cpp — this repository's drop-in replacement
//                        of the reference's entry-point file — linked against libdmpc.so, so that the R-facing shim is
//                        exercised end to end (fake R objects -> Rcpp-style marshalling -> C-ABI -> GPU) without R.
#include <RcppArmadillo.h>
#include <RcppGSL.h>
#include <cstdlib>
#include <cstring>

#ifdef HARNESS_DMPC
#include <Rcpp.h>
#include "dmpc.h"
#else
#include "AugTree.h"
#endif

namespace Rcpp { SEXPREC g_shim_nil(SEXPREC::NIL); }

using namespace Rcpp;
using namespace arma;

// prototypes of the reference's exported functions (definitions: clusterFunArmadillo.cpp)
List logLikCpp(IntegerMatrix& edgeMat, NumericVector& clusterMRCAs, NumericVector& limProbsVec,
               List& withinTransMatList, List& betweenTransMatList, int numOpenMP, List alignmentBin,
               uint numTips, uint numLoci, uint withinMatListIndex, uint betweenMatListIndex);
SEXP getConvertedAlignment(SEXP& equivVector, CharacterMatrix& alignmentAlphaMat);
double newBetweenTransProbsLogLik(SEXP AugTreePointer, List& withinTransProbs, List& newBetweenTransProbs,
                                  uint& numOpenMP, uint& newBetweenMatListIndex, NumericVector& limProbs);
double newWithinTransProbsLogLik(SEXP AugTreePointer, List& newWithinTransProbs, List& betweenTransProbs,
                                 uint& numOpenMP, uint& newWithinMatListIndex, NumericVector& limProbs);
List withinClusNNIlogLik(SEXP AugTreePointer, List& withinTransProbs, List& betweenTransProbs,
                         uint& MRCAofClusForNNI, uint& numMovesNNI, uint& numOpenMP, NumericVector& limProbs);
List betweenClusNNIlogLik(SEXP AugTreePointer, List& withinTransProbs, List& betweenTransProbs,
                          NumericVector& clusterMRCAs, uint& numMovesNNI, uint& numOpenMP, NumericVector& limProbs);
double clusSplitMergeLogLik(SEXP AugTreePointer, List& withinTransProbs, List& betweenTransProbs,
                            IntegerVector& clusMRCAsToSplitOrMerge, uint& numOpenMP, NumericVector& limProbs);
void RestorePreviousConfig(SEXP AugTreePointer, IntegerMatrix& edgeMat, int& withinMatListIndex,
                           int& betweenMatListIndex, bool NNImove, IntegerVector& clusterMRCAs, bool splitMergeMove);
void finalDeallocate(SEXP AugTreePointer);
double getMaxMapSizeEstimate(double& allowedSizeInMegs, unsigned int& numRateCats);
void checkAndCullMap(SEXP AugTreePointer, unsigned int& allowedNumberOfElements, float& cullProportion,
                     List& withinTransProbs, List& betweenTransProbs, NumericVector& limProbs);

namespace {
struct Scope {   // frees every shim record allocated during one call
  size_t mark;
  Scope() : mark(shim::arena().size()) {}
  ~Scope() { shim::release_from(mark); }
};
SEXP realVec(const double* x, int n) { SEXP s = shim::alloc(SEXPREC::REAL); s->real.assign(x, x + n); return s; }
SEXP intVec(const int* x, int n) { SEXP s = shim::alloc(SEXPREC::INT); s->integer.assign(x, x + n); return s; }
SEXP intMat(const int* x, int nr, int nc) { SEXP s = intVec(x, nr * nc); s->nrow = nr; s->ncol = nc; return s; }
SEXP matList(const double* m, int R) {   // R column-major 4x4 matrices
  SEXP l = shim::alloc(SEXPREC::LIST);
  for (int r = 0; r < R; r++) { SEXP x = realVec(m + 16 * r, 16); x->nrow = 4; x->ncol = 4; l->list.push_back(x); }
  return l;
}
#ifdef HARNESS_DMPC
typedef dmpc_tree TreeT;
#else
typedef AugTree TreeT;
#endif
struct Handle {
  TreeT* tree;
  SEXPREC ext;   // persistent external-pointer record handed to the entry points on every call
  int numRates, numTips, numLoci;
  Handle() : tree(0), ext(SEXPREC::EXTPTR), numRates(0), numTips(0), numLoci(0) {}
};
void copyEdge(SEXP e, int* out) { for (size_t i = 0; i < e->real.size(); i++) out[i] = (int)e->real[i]; }
std::string g_err;
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// getConvertedAlignment (clusterFunArmadillo.cpp:97-113) on a flat char matrix [numTips][numLoci]
// (row-major); result packed as one 4-bit mask per cell.  Returns the number of cells for which the
// reference produced an EMPTY indicator (unknown symbol).
long ref_convert_alignment(const char* chars, int numTips, int numLoci, const char* states, unsigned char* out) {
  Scope sc;
  SEXP eq = shim::alloc(SEXPREC::STR);
  for (int j = 0; j < 4; j++) eq->str.push_back(std::string(1, states[j]));
  SEXP cm = shim::alloc(SEXPREC::STR);
  cm->nrow = numTips; cm->ncol = numLoci;
  cm->str.resize((size_t)numTips * numLoci);
  for (int i = 0; i < numTips; i++)
    for (int j = 0; j < numLoci; j++) cm->str[(size_t)i + (size_t)j * numTips] = std::string(1, chars[(size_t)i * numLoci + j]);
  CharacterMatrix M(cm);
  SEXP res = getConvertedAlignment(eq, M);
  long bad = 0;
  for (int i = 0; i < numTips; i++)
    for (int j = 0; j < numLoci; j++) {
      SEXP cell = res->list[i]->list[j];
      unsigned char m = 0;
      if (cell->real.size() != 4) bad++;
      else for (int k = 0; k < 4; k++) if (cell->real[k] != 0.0) m |= (unsigned char)(1u << k);
      out[(size_t)i * numLoci + j] = m;
    }
  return bad;
}

void* ref_loglik_create(const int* edge, int nEdgeRows, const double* clusterMRCAs, int nMRCAs,
                        const double* limProbs, const double* within, const double* between, int numRates,
                        const unsigned char* tipmask, int numTips, int numLoci, int withinIdx, int betweenIdx,
                        double* logLik) {
  Scope sc;
  try {
    IntegerMatrix e(intMat(edge, nEdgeRows, 2));
    NumericVector m(realVec(clusterMRCAs, nMRCAs)), pi(realVec(limProbs, 4));
    List w(matList(within, numRates)), b(matList(between, numRates));
    // alignmentBin: list (tips) of lists (loci) of 0/1 vectors, as getConvertedAlignment returns it
    SEXP al = shim::alloc(SEXPREC::LIST);
    for (int i = 0; i < numTips; i++) {
      SEXP row = shim::alloc(SEXPREC::LIST);
      row->list.reserve(numLoci);
      for (int j = 0; j < numLoci; j++) {
        double x[4];
        unsigned char mk = tipmask[(size_t)i * numLoci + j];
        for (int k = 0; k < 4; k++) x[k] = (mk >> k) & 1 ? 1.0 : 0.0;
        row->list.push_back(realVec(x, 4));
      }
      al->list.push_back(row);
    }
#ifdef HARNESS_DMPC
    if (getenv("SHIM_FAST_MASKS")) {     // what getConvertedAlignment of the drop-in shim attaches: packed masks, column-major
      SEXP raw = shim::alloc(SEXPREC::RAW);
      raw->raw.resize((size_t)numTips * numLoci);
      for (int i = 0; i < numTips; i++)
        for (int j = 0; j < numLoci; j++) raw->raw[(size_t)i + (size_t)j * numTips] = tipmask[(size_t)i * numLoci + j];
      al->attrs["dmpc_masks"] = raw;
    }
#endif
    List res = logLikCpp(e, m, pi, w, b, 1, List(al), (uint)numTips, (uint)numLoci, (uint)withinIdx, (uint)betweenIdx);
    Handle* h = new Handle;
    h->tree = static_cast<TreeT*>(res[1]->ptr);
    h->ext.ptr = h->tree;
    h->numRates = numRates; h->numTips = numTips; h->numLoci = numLoci;
    *logLik = res[0]->real[0];
    return h;
  } catch (std::exception& ex) { g_err = ex.what(); return 0; }
}

double ref_new_between(void* hv, const double* within, const double* newBetween, int numRates, int newIdx, const double* limProbs) {
  Scope sc; Handle* h = (Handle*)hv;
  List w(matList(within, numRates)), b(matList(newBetween, numRates));
  NumericVector pi(realVec(limProbs, 4));
  uint omp = 1, idx = (uint)newIdx;
  return newBetweenTransProbsLogLik(&h->ext, w, b, omp, idx, pi);
}

double ref_new_within(void* hv, const double* newWithin, const double* between, int numRates, int newIdx, const double* limProbs) {
  Scope sc; Handle* h = (Handle*)hv;
  List w(matList(newWithin, numRates)), b(matList(between, numRates));
  NumericVector pi(realVec(limProbs, 4));
  uint omp = 1, idx = (uint)newIdx;
  return newWithinTransProbsLogLik(&h->ext, w, b, omp, idx, pi);
}

double ref_within_nni(void* hv, const double* within, const double* between, int numRates, int mrca, int numMoves,
                      const double* limProbs, int* edgeOut) {
  Scope sc; Handle* h = (Handle*)hv;
  List w(matList(within, numRates)), b(matList(between, numRates));
  NumericVector pi(realVec(limProbs, 4));
  uint omp = 1, m = (uint)mrca, nm = (uint)numMoves;
  List res = withinClusNNIlogLik(&h->ext, w, b, m, nm, omp, pi);
  copyEdge(res[1], edgeOut);
  return res[0]->real[0];
}

double ref_between_nni(void* hv, const double* within, const double* between, int numRates, const double* clusterMRCAs,
                       int nMRCAs, int numMoves, const double* limProbs, int* edgeOut) {
  Scope sc; Handle* h = (Handle*)hv;
  List w(matList(within, numRates)), b(matList(between, numRates));
  NumericVector pi(realVec(limProbs, 4)), mr(realVec(clusterMRCAs, nMRCAs));
  uint omp = 1, nm = (uint)numMoves;
  List res = betweenClusNNIlogLik(&h->ext, w, b, mr, nm, omp, pi);
  copyEdge(res[1], edgeOut);
  return res[0]->real[0];
}

double ref_split_merge(void* hv, const double* within, const double* between, int numRates, const int* mrcas, int n,
                       const double* limProbs) {
  Scope sc; Handle* h = (Handle*)hv;
  List w(matList(within, numRates)), b(matList(between, numRates));
  NumericVector pi(realVec(limProbs, 4));
  IntegerVector mr(intVec(mrcas, n));
  uint omp = 1;
  return clusSplitMergeLogLik(&h->ext, w, b, mr, omp, pi);
}

void ref_restore(void* hv, const int* edge, int nEdgeRows, int withinIdx, int betweenIdx, int nniMove,
                 const int* clusterMRCAs, int nMRCAs, int splitMergeMove) {
  Scope sc; Handle* h = (Handle*)hv;
  IntegerMatrix e(intMat(edge, nEdgeRows, 2));
  IntegerVector mr(intVec(clusterMRCAs, nMRCAs));
  RestorePreviousConfig(&h->ext, e, withinIdx, betweenIdx, nniMove != 0, mr, splitMergeMove != 0);
}

void ref_final_deallocate(void* hv) {
  Handle* h = (Handle*)hv;
  finalDeallocate(&h->ext);
#ifndef HARNESS_DMPC
  delete h->tree;   // the reference leaks the AugTree itself (clusterFunArmadillo.cpp:288-294); the harness does not
#endif
  delete h;
}

double ref_max_map_size_estimate(double mb, unsigned numRateCats) { return getMaxMapSizeEstimate(mb, numRateCats); }

void ref_check_and_cull_map(void* hv, unsigned allowed, float cullProportion, const double* within, const double* between,
                            int numRates, const double* limProbs) {
  Scope sc; Handle* h = (Handle*)hv;
  List w(matList(within, numRates)), b(matList(between, numRates));
  NumericVector pi(realVec(limProbs, 4));
  checkAndCullMap(&h->ext, allowed, cullProportion, w, b, pi);
}

#ifdef HARNESS_DMPC
// ---- the SEXP-level glue of the R package skeleton (rpkg/src/RcppExports.cpp), compiled against the stand-in too
extern "C" SEXP _DMphyClus_getMaxMapSizeEstimate(SEXP, SEXP);
extern "C" void R_init_DMphyClus(DllInfo*);
// one call through the registered SEXP entry point (host-only arithmetic: needs no GPU)
double ref_glue_max_map_size_estimate(double mb, unsigned numRateCats) {
  Scope sc;
  SEXP r = _DMphyClus_getMaxMapSizeEstimate(wrap(mb), wrap((double)numRateCats));
  return r->real.empty() ? -1.0 : r->real[0];
}
// the routine table R_init_DMphyClus registers: names separated by ';' and the argument counts
int ref_registered_routines(char* names, int cap, int* nargs, int maxRoutines) {
  DllInfo d; d.calls = 0;
  R_init_DMphyClus(&d);
  std::string all;
  int n = 0;
  for (const R_CallMethodDef* c = d.calls; c && c->name; c++) {
    if (n < maxRoutines) nargs[n] = c->numArgs;
    all += c->name; all += ';';
    n++;
  }
  std::strncpy(names, all.c_str(), cap - 1); names[cap - 1] = 0;
  return n;
}
// ---- introspection through the C-ABI
double ref_loglik(void*) { return 0.0; }
long ref_dictionary_size(void*) { return 0; }
void ref_get_partial(void* hv, int node, double* sol, float* expo) { dmpc_get_partial(((Handle*)hv)->tree, node, sol, expo); }
void ref_get_topology(void* hv, int* parent, int* child0, int* child1, int* within) {
  dmpc_get_topology(((Handle*)hv)->tree, parent, child0, child1, within);
}
#else
// ---- introspection for bit-level comparison with oracle/dmpc_oracle.cpp
double ref_loglik(void* hv) { return ((Handle*)hv)->tree->GetLoglik(); }
long ref_dictionary_size(void* hv) { return (long)((Handle*)hv)->tree->GetSolutionDictionary()->size(); }
void ref_get_partial(void* hv, int node, double* sol /*[S][R][4]*/, float* expo /*[S][R]*/) {
  Handle* h = (Handle*)hv;
  TreeNode* v = h->tree->GetVertexVector().at(node);
  uint S = h->tree->GetNumLoci(), R = h->tree->GetNumRateCats();
  for (uint l = 0; l < S; l++)
    for (uint r = 0; r < R; r++) {
      vec s = v->GetSolution(l, r);
      for (int k = 0; k < 4; k++) sol[((size_t)l * R + r) * 4 + k] = s[k];
      expo[(size_t)l * R + r] = v->GetExponent(l, r);
    }
}
void ref_get_topology(void* hv, int* parent, int* child0, int* child1, int* within) {
  Handle* h = (Handle*)hv;
  std::vector<TreeNode*> vv = h->tree->GetVertexVector();
  for (size_t i = 0; i < vv.size(); i++) {
    TreeNode* n = vv[i];
    parent[i] = n->GetParent() ? (int)n->GetParent()->GetId() : -1;
    std::vector<TreeNode*> ch = n->GetChildren();
    bool tip = ch.at(0) == NULL;
    child0[i] = tip ? -1 : (int)ch.at(0)->GetId();
    child1[i] = tip ? -1 : (int)ch.at(1)->GetId();
    within[i] = n->GetWithinParentBranch() ? 1 : 0;
  }
}

#endif

}  // extern "C"
