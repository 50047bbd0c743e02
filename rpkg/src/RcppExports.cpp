// RcppExports.cpp — SEXP-level glue of the eleven entry points and their registration with R.
//
// Takes the place of the file Rcpp::compileAttributes() generates for the upstream package
// (/root/reference/src/RcppExports.cpp:10-197): the same eleven `_DMphyClus_*` symbols with the same argument counts
// (its routine table, :179-192), registered by R_init_DMphyClus, each one converting its SEXP arguments, opening an
// RNGScope, calling the C++ function of clusterFunArmadillo.cpp and wrapping the result inside BEGIN_RCPP / END_RCPP
// (C++ exceptions become R errors).  Written once by hand as ONE variadic template instead of eleven generated bodies;
// running Rcpp::compileAttributes() on the assembled package overwrites it with an equivalent generated file.
#include <Rcpp.h>

using namespace Rcpp;
typedef unsigned int uint;

// the exported C++ functions (clusterFunArmadillo.cpp in this directory; signatures = upstream's :36-305)
List logLikCpp(IntegerMatrix& edgeMat, NumericVector& clusterMRCAs, NumericVector& limProbsVec, List& withinTransMatList,
               List& betweenTransMatList, int numOpenMP, List alignmentBin, uint numTips, uint numLoci,
               uint withinMatListIndex, uint betweenMatListIndex);
SEXP getConvertedAlignment(SEXP& equivVector, CharacterMatrix& alignmentAlphaMat);
double newBetweenTransProbsLogLik(SEXP AugTreePointer, List& withinTransProbs, List& newBetweenTransProbs, uint& numOpenMP,
                                  uint& newBetweenMatListIndex, NumericVector& limProbs);
double newWithinTransProbsLogLik(SEXP AugTreePointer, List& newWithinTransProbs, List& betweenTransProbs, uint& numOpenMP,
                                 uint& newWithinMatListIndex, NumericVector& limProbs);
List withinClusNNIlogLik(SEXP AugTreePointer, List& withinTransProbs, List& betweenTransProbs, uint& MRCAofClusForNNI,
                         uint& numMovesNNI, uint& numOpenMP, NumericVector& limProbs);
List betweenClusNNIlogLik(SEXP AugTreePointer, List& withinTransProbs, List& betweenTransProbs, NumericVector& clusterMRCAs,
                          uint& numMovesNNI, uint& numOpenMP, NumericVector& limProbs);
double clusSplitMergeLogLik(SEXP AugTreePointer, List& withinTransProbs, List& betweenTransProbs,
                            IntegerVector& clusMRCAsToSplitOrMerge, uint& numOpenMP, NumericVector& limProbs);
void RestorePreviousConfig(SEXP AugTreePointer, IntegerMatrix& edgeMat, int& withinMatListIndex, int& betweenMatListIndex,
                           bool NNImove, IntegerVector& clusterMRCAs, bool splitMergeMove);
void finalDeallocate(SEXP AugTreePointer);
double getMaxMapSizeEstimate(double& allowedSizeInMegs, unsigned& numRateCats);
void checkAndCullMap(SEXP AugTreePointer, unsigned& allowedNumberOfElements, float& cullProportion, List& withinTransProbs,
                     List& betweenTransProbs, NumericVector& limProbs);

namespace {

// one converted argument: Rcpp's own input_parameter machinery (the same type the generated glue instantiates)
template <class A>
using Arg = typename Rcpp::traits::input_parameter<A>::type;

template <class R, class... A>
struct Glue {
  template <R (*F)(A...)>
  static SEXP call(typename std::conditional<true, SEXP, A>::type... s) {
    BEGIN_RCPP
    Rcpp::RNGScope scope;
    return invoke<F>(Arg<A>(s)...);
    END_RCPP
  }
  template <R (*F)(A...), class... H>
  static SEXP invoke(H&&... holders) { return Rcpp::wrap(F(holders...)); }
};
template <class... A>
struct Glue<void, A...> {
  template <void (*F)(A...)>
  static SEXP call(typename std::conditional<true, SEXP, A>::type... s) {
    BEGIN_RCPP
    Rcpp::RNGScope scope;
    invoke<F>(Arg<A>(s)...);
    return R_NilValue;
    END_RCPP
  }
  template <void (*F)(A...), class... H>
  static void invoke(H&&... holders) { F(holders...); }
};
template <class R, class... A>
Glue<R, A...> glueOf(R (*)(A...));
#define GLUE(fn) decltype(glueOf(&fn))::call<&fn>

}  // namespace

#define DMPC_EXPORT1(fn) RcppExport SEXP _DMphyClus_##fn(SEXP a) { return GLUE(fn)(a); }
#define DMPC_EXPORT2(fn) RcppExport SEXP _DMphyClus_##fn(SEXP a, SEXP b) { return GLUE(fn)(a, b); }
#define DMPC_EXPORT6(fn) RcppExport SEXP _DMphyClus_##fn(SEXP a, SEXP b, SEXP c, SEXP d, SEXP e, SEXP f) { return GLUE(fn)(a, b, c, d, e, f); }
#define DMPC_EXPORT7(fn) RcppExport SEXP _DMphyClus_##fn(SEXP a, SEXP b, SEXP c, SEXP d, SEXP e, SEXP f, SEXP g) { return GLUE(fn)(a, b, c, d, e, f, g); }
#define DMPC_EXPORT11(fn) \
  RcppExport SEXP _DMphyClus_##fn(SEXP a, SEXP b, SEXP c, SEXP d, SEXP e, SEXP f, SEXP g, SEXP h, SEXP i, SEXP j, SEXP k) { \
    return GLUE(fn)(a, b, c, d, e, f, g, h, i, j, k); }

DMPC_EXPORT11(logLikCpp)
DMPC_EXPORT2(getConvertedAlignment)
DMPC_EXPORT6(newBetweenTransProbsLogLik)
DMPC_EXPORT6(newWithinTransProbsLogLik)
DMPC_EXPORT7(withinClusNNIlogLik)
DMPC_EXPORT7(betweenClusNNIlogLik)
DMPC_EXPORT6(clusSplitMergeLogLik)
DMPC_EXPORT7(RestorePreviousConfig)
DMPC_EXPORT1(finalDeallocate)
DMPC_EXPORT2(getMaxMapSizeEstimate)
DMPC_EXPORT6(checkAndCullMap)

#define DMPC_ROUTINE(fn, n) {"_DMphyClus_" #fn, (DL_FUNC)&_DMphyClus_##fn, n}
static const R_CallMethodDef dmpcRoutines[] = {
    DMPC_ROUTINE(logLikCpp, 11),
    DMPC_ROUTINE(getConvertedAlignment, 2),
    DMPC_ROUTINE(newBetweenTransProbsLogLik, 6),
    DMPC_ROUTINE(newWithinTransProbsLogLik, 6),
    DMPC_ROUTINE(withinClusNNIlogLik, 7),
    DMPC_ROUTINE(betweenClusNNIlogLik, 7),
    DMPC_ROUTINE(clusSplitMergeLogLik, 6),
    DMPC_ROUTINE(RestorePreviousConfig, 7),
    DMPC_ROUTINE(finalDeallocate, 1),
    DMPC_ROUTINE(getMaxMapSizeEstimate, 2),
    DMPC_ROUTINE(checkAndCullMap, 6),
    {NULL, NULL, 0}};

RcppExport void R_init_DMphyClus(DllInfo* dll) {
  R_registerRoutines(dll, NULL, dmpcRoutines, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
