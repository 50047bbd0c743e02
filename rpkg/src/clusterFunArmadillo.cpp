// clusterFunArmadillo.cpp — drop-in replacement of the reference's src/clusterFunArmadillo.cpp.
//
// Same eleven [[Rcpp::export]] functions, same names, argument lists and return shapes
// (/root/reference/src/clusterFunArmadillo.cpp:36-332), so R/RcppExports.R, src/RcppExports.cpp (regenerate with
// Rcpp::compileAttributes()) and every R caller stay untouched.  Each body only marshals R objects into plain
// arrays and calls the C-ABI of libdmpc.so (include/dmpc.h), where the likelihood runs on the GPU.  The
// reference's AugTree.*, IntermediateNode.*, InputNode.*, TreeNode.h and helper.h are no longer compiled, and the
// package no longer links Armadillo, GSL or Boost (see Makevars next to this file).
//
// Not compiled in this repository's image (no R): INTEGRATION.md says how to build it on a box with R.
#include <Rcpp.h>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "dmpc.h"

using namespace Rcpp;
typedef unsigned int uint;      // the reference's signatures spell it this way (glibc's <sys/types.h> typedef; restated for other libcs)

namespace {

[[noreturn]] void fail(int rc) { stop("libdmpc (status %d): %s", rc, dmpc_last_error()); }
inline void check(int rc) { if (rc != DMPC_OK) fail(rc); }

// list (one entry per rate category) of 4x4 numeric matrices -> R*16 doubles, each matrix column-major as R stores it
std::vector<double> matList(const List& l) {
  std::vector<double> out;
  out.reserve(16 * l.size());
  for (int r = 0; r < l.size(); r++) {
    NumericVector m(l[r]);
    if (m.size() != 16) stop("transition matrices must be 4 x 4");
    for (int i = 0; i < 16; i++) out.push_back(m[i]);
  }
  return out;
}
std::vector<double> numVec(NumericVector& v) {
  std::vector<double> out(v.size());
  for (int i = 0; i < v.size(); i++) out[i] = v[i];
  return out;
}
std::vector<int32_t> intVec(IntegerVector& v) {
  std::vector<int32_t> out(v.size());
  for (int i = 0; i < v.size(); i++) out[i] = v[i];
  return out;
}
std::vector<int32_t> edgeVec(IntegerMatrix& e) {        // column-major already: parents, then children
  std::vector<int32_t> out((size_t)e.nrow() * e.ncol());
  IntegerVector flat(static_cast<SEXP>(e));
  for (size_t i = 0; i < out.size(); i++) out[i] = flat[(int)i];
  return out;
}
dmpc_tree* handle(SEXP p) {
  if (p == NULL) throw ::Rcpp::exception("pointer is null.");          // clusterFunArmadillo.cpp:136
  XPtr<dmpc_tree> x(p);
  return x.get();
}
NumericVector edgeOut(const std::vector<int32_t>& e, int nRows) {        // umat (2N-2) x 2, as wrap(arma::umat) gives R
  NumericVector out(e.begin(), e.end());
  out.attr("dim") = IntegerVector::create(nRows, 2);
  return out;
}
dmpc_options optionsFromEnv() {       // the R signature has no room for options (SURVEY.md §5 "Config / flags")
  dmpc_options o;
  dmpc_default_options(&o);
  if (const char* d = std::getenv("DMPC_DEVICE")) o.device = std::atoi(d);
  if (const char* q = std::getenv("DMPC_QUIRK_CARRY")) o.quirk_carry = std::atoi(q);
  // DMPC_DEVICES=0,1,...,7: single-process multi-GPU — one handle whose sites are sharded over these devices; R stays
  // one process and DMphyClusChain's signature (R/DMphyClusInterface.R:71) stays what it is
  if (const char* l = std::getenv("DMPC_DEVICES")) {
    int n = 0;
    const char* p = l;
    while (*p && n < 8) {
      char* end = NULL;
      const long d = std::strtol(p, &end, 10);
      if (end == p) break;
      o.devices[n++] = (int)d;
      p = end;
      while (*p == ',' || *p == ' ') p++;
    }
    if (n > 1) o.n_devices = n; else if (n == 1) o.device = o.devices[0];
  }
  return o;
}

}  // namespace

// [[Rcpp::export]]
List logLikCpp(IntegerMatrix& edgeMat, NumericVector& clusterMRCAs, NumericVector& limProbsVec, List& withinTransMatList,
               List& betweenTransMatList, int numOpenMP, List alignmentBin, uint numTips, uint numLoci,
               uint withinMatListIndex, uint betweenMatListIndex) {
  // alignmentBin is getConvertedAlignment's value: list(numTips) of list(numLoci) of 0/1 vectors; the packed masks
  // ride along as attribute "dmpc_masks" (raw numTips x numLoci, column-major) when it came from this library.
  std::vector<uint8_t> masks((size_t)numTips * numLoci);
  RObject packed(alignmentBin.attr("dmpc_masks"));
  auto maskOfCell = [&](uint i, uint j) -> uint8_t {
    List row(alignmentBin[i]);
    IntegerVector ind(row[j]);
    uint8_t m = 0;
    for (int k = 0; k < ind.size() && k < 4; k++) m |= (uint8_t)((ind[k] != 0) << k);
    return m;
  };
  // The attribute is trusted only when it still describes THIS list: raw, numTips x numLoci, and equal to the list on a
  // spread of probe cells (R's `[[<-` keeps attributes, so an edited alignmentBin could otherwise be evaluated with
  // stale masks).  Anything else falls back to the list itself.
  bool usePacked = false;
  if (!packed.isNULL() && TYPEOF(packed) == RAWSXP && (size_t)alignmentBin.size() == (size_t)numTips) {
    RawVector raw(packed);
    RObject dim(raw.attr("dim"));
    if (!dim.isNULL()) {
      IntegerVector d(dim);
      usePacked = d.size() == 2 && (uint)d[0] == numTips && (uint)d[1] == numLoci && (size_t)raw.size() == (size_t)numTips * numLoci;
    }
    const size_t cells = (size_t)numTips * numLoci, probes = cells < 257 ? cells : 257;
    for (size_t q = 0; usePacked && q < probes; q++) {
      const size_t c = probes == cells ? q : (q * (cells - 1)) / (probes - 1);
      const uint i = (uint)(c / numLoci), j = (uint)(c % numLoci);
      if (List(alignmentBin[i]).size() != (int)numLoci || raw[(size_t)i + (size_t)j * numTips] != maskOfCell(i, j)) usePacked = false;
    }
    if (usePacked)
      for (uint i = 0; i < numTips; i++)
        for (uint j = 0; j < numLoci; j++) masks[(size_t)i * numLoci + j] = raw[(size_t)i + (size_t)j * numTips];
  }
  if (!usePacked) {
    for (uint i = 0; i < numTips; i++) {
      List row(alignmentBin[i]);
      for (uint j = 0; j < numLoci; j++) {
        IntegerVector ind(row[j]);
        uint8_t m = 0;
        for (int k = 0; k < ind.size() && k < 4; k++) m |= (uint8_t)((ind[k] != 0) << k);
        masks[(size_t)i * numLoci + j] = m;
      }
    }
  }
  std::vector<int32_t> edge = edgeVec(edgeMat);
  std::vector<double> mrca = numVec(clusterMRCAs), pi = numVec(limProbsVec), w = matList(withinTransMatList),
                      b = matList(betweenTransMatList);
  dmpc_options o = optionsFromEnv();
  dmpc_tree* t = NULL;
  double ll = 0;
  check(dmpc_loglik_create(edge.data(), edgeMat.nrow(), mrca.data(), (int)mrca.size(), pi.data(), w.data(), b.data(),
                           withinTransMatList.size(), masks.data(), (int)numTips, (int)numLoci, (int)withinMatListIndex,
                           (int)betweenMatListIndex, &o, &t, &ll));
  XPtr<dmpc_tree> p(t, false);                                   // freed by finalDeallocate, as in the reference
  return List::create(Named("logLik") = ll, Named("solutionPointer") = p, Named("alignmentBinPointer") = R_NilValue);
}

// [[Rcpp::export]]
SEXP getConvertedAlignment(SEXP& equivVector, CharacterMatrix& alignmentAlphaMat) {
  std::vector<std::string> states = as<std::vector<std::string> >(equivVector);
  if (states.size() != 4) stop("4-state DNA only");
  char st[4];
  for (int j = 0; j < 4; j++) {
    if (states[j].size() != 1) stop("state labels must be single characters");
    st[j] = states[j][0];
  }
  const int n = alignmentAlphaMat.nrow(), s = alignmentAlphaMat.ncol();
  std::vector<char> chars((size_t)n * s);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < s; j++) {
      // a multi-character cell matches no key of the reference's string-keyed map (clusterFunArmadillo.cpp:64-93): unknown
      const std::string cell(alignmentAlphaMat(i, j));
      chars[(size_t)i * s + j] = cell.size() == 1 ? cell[0] : '?';
    }
  std::vector<uint8_t> masks(chars.size());
  dmpc_convert_alignment(chars.data(), (long)chars.size(), st, masks.data());
  // R-visible shape of the reference (list of lists of 0/1 vectors; an unknown symbol gives an empty vector)
  List out(n);
  RawVector raw((size_t)n * s);
  for (int i = 0; i < n; i++) {
    List row(s);
    for (int j = 0; j < s; j++) {
      const uint8_t m = masks[(size_t)i * s + j];
      raw[(size_t)i + (size_t)j * n] = m;
      if (m == 0) { row[j] = NumericVector(0); continue; }
      row[j] = NumericVector::create(m & 1, (m >> 1) & 1, (m >> 2) & 1, (m >> 3) & 1);
    }
    out[i] = row;
  }
  raw.attr("dim") = IntegerVector::create(n, s);
  out.attr("dmpc_masks") = raw;
  return out;
}

// [[Rcpp::export]]
double newBetweenTransProbsLogLik(SEXP AugTreePointer, List& withinTransProbs, List& newBetweenTransProbs, uint& numOpenMP,
                                  uint& newBetweenMatListIndex, NumericVector& limProbs) {
  std::vector<double> w = matList(withinTransProbs), b = matList(newBetweenTransProbs), pi = numVec(limProbs);
  double ll = 0;
  check(dmpc_new_between_transprobs_loglik(handle(AugTreePointer), w.data(), b.data(), (int)newBetweenMatListIndex, pi.data(), &ll));
  return ll;
}

// [[Rcpp::export]]
double newWithinTransProbsLogLik(SEXP AugTreePointer, List& newWithinTransProbs, List& betweenTransProbs, uint& numOpenMP,
                                 uint& newWithinMatListIndex, NumericVector& limProbs) {
  std::vector<double> w = matList(newWithinTransProbs), b = matList(betweenTransProbs), pi = numVec(limProbs);
  double ll = 0;
  check(dmpc_new_within_transprobs_loglik(handle(AugTreePointer), w.data(), b.data(), (int)newWithinMatListIndex, pi.data(), &ll));
  return ll;
}

// [[Rcpp::export]]
List withinClusNNIlogLik(SEXP AugTreePointer, List& withinTransProbs, List& betweenTransProbs, uint& MRCAofClusForNNI,
                         uint& numMovesNNI, uint& numOpenMP, NumericVector& limProbs) {
  std::vector<double> w = matList(withinTransProbs), b = matList(betweenTransProbs), pi = numVec(limProbs);
  dmpc_tree* t = handle(AugTreePointer);
  dmpc_stats st;
  check(dmpc_get_stats(t, &st));
  const int nRows = 2 * st.num_tips - 2;
  std::vector<int32_t> edge(2 * (size_t)nRows);
  double ll = 0;
  check(dmpc_within_clus_nni_loglik(t, w.data(), b.data(), (int)MRCAofClusForNNI, (int)numMovesNNI, pi.data(), &ll, edge.data()));
  return List::create(Named("logLik") = ll, Named("edge") = edgeOut(edge, nRows));
}

// [[Rcpp::export]]
List betweenClusNNIlogLik(SEXP AugTreePointer, List& withinTransProbs, List& betweenTransProbs, NumericVector& clusterMRCAs,
                          uint& numMovesNNI, uint& numOpenMP, NumericVector& limProbs) {
  std::vector<double> w = matList(withinTransProbs), b = matList(betweenTransProbs), pi = numVec(limProbs),
                      m = numVec(clusterMRCAs);
  dmpc_tree* t = handle(AugTreePointer);
  dmpc_stats st;
  check(dmpc_get_stats(t, &st));
  const int nRows = 2 * st.num_tips - 2;
  std::vector<int32_t> edge(2 * (size_t)nRows);
  double ll = 0;
  check(dmpc_between_clus_nni_loglik(t, w.data(), b.data(), m.data(), (int)m.size(), (int)numMovesNNI, pi.data(), &ll, edge.data()));
  return List::create(Named("logLik") = ll, Named("edge") = edgeOut(edge, nRows));
}

// [[Rcpp::export]]
double clusSplitMergeLogLik(SEXP AugTreePointer, List& withinTransProbs, List& betweenTransProbs,
                            IntegerVector& clusMRCAsToSplitOrMerge, uint& numOpenMP, NumericVector& limProbs) {
  std::vector<double> w = matList(withinTransProbs), b = matList(betweenTransProbs), pi = numVec(limProbs);
  std::vector<int32_t> m = intVec(clusMRCAsToSplitOrMerge);
  double ll = 0;
  check(dmpc_clus_split_merge_loglik(handle(AugTreePointer), w.data(), b.data(), m.data(), (int)m.size(), pi.data(), &ll));
  return ll;
}

// [[Rcpp::export]]
void RestorePreviousConfig(SEXP AugTreePointer, IntegerMatrix& edgeMat, int& withinMatListIndex, int& betweenMatListIndex,
                           bool NNImove, IntegerVector& clusterMRCAs, bool splitMergeMove) {
  std::vector<int32_t> edge = edgeVec(edgeMat), m = intVec(clusterMRCAs);
  check(dmpc_restore_previous_config(handle(AugTreePointer), edge.data(), edgeMat.nrow(), withinMatListIndex,
                                     betweenMatListIndex, NNImove, m.data(), (int)m.size(), splitMergeMove));
}

// [[Rcpp::export]]
void finalDeallocate(SEXP AugTreePointer) { check(dmpc_final_deallocate(handle(AugTreePointer))); }

// [[Rcpp::export]]
double getMaxMapSizeEstimate(double& allowedSizeInMegs, unsigned int& numRateCats) {
  return dmpc_get_max_map_size_estimate(allowedSizeInMegs, numRateCats);
}

// [[Rcpp::export]]
void checkAndCullMap(SEXP AugTreePointer, unsigned int& allowedNumberOfElements, float& cullProportion, List& withinTransProbs,
                     List& betweenTransProbs, NumericVector& limProbs) {
  std::vector<double> w = matList(withinTransProbs), b = matList(betweenTransProbs), pi = numVec(limProbs);
  check(dmpc_check_and_cull_map(handle(AugTreePointer), allowedNumberOfElements, cullProportion, w.data(), b.data(), pi.data()));
}
