# R-side stubs of the eleven compiled entry points (what Rcpp::compileAttributes() writes for the upstream package,
# /root/reference/R/RcppExports.R:4-46), built from ONE table: the entry point's name and its formal arguments, in the
# order the compiled routine takes them.  The upstream R code calls these by name with named arguments
# (R/DMphyClusSource.R:90,235,239,337,348,391,409; R/DMphyClusInterface.R:224,251,252), so names and order are the
# contract.  `Rcpp::compileAttributes()` on the assembled package regenerates an equivalent file.
.dmpcEntryPoints <- list(
  logLikCpp = c("edgeMat", "clusterMRCAs", "limProbsVec", "withinTransMatList", "betweenTransMatList", "numOpenMP",
                "alignmentBin", "numTips", "numLoci", "withinMatListIndex", "betweenMatListIndex"),
  getConvertedAlignment = c("equivVector", "alignmentAlphaMat"),
  newBetweenTransProbsLogLik = c("AugTreePointer", "withinTransProbs", "newBetweenTransProbs", "numOpenMP",
                                 "newBetweenMatListIndex", "limProbs"),
  newWithinTransProbsLogLik = c("AugTreePointer", "newWithinTransProbs", "betweenTransProbs", "numOpenMP",
                                "newWithinMatListIndex", "limProbs"),
  withinClusNNIlogLik = c("AugTreePointer", "withinTransProbs", "betweenTransProbs", "MRCAofClusForNNI", "numMovesNNI",
                          "numOpenMP", "limProbs"),
  betweenClusNNIlogLik = c("AugTreePointer", "withinTransProbs", "betweenTransProbs", "clusterMRCAs", "numMovesNNI",
                           "numOpenMP", "limProbs"),
  clusSplitMergeLogLik = c("AugTreePointer", "withinTransProbs", "betweenTransProbs", "clusMRCAsToSplitOrMerge",
                           "numOpenMP", "limProbs"),
  RestorePreviousConfig = c("AugTreePointer", "edgeMat", "withinMatListIndex", "betweenMatListIndex", "NNImove",
                            "clusterMRCAs", "splitMergeMove"),
  finalDeallocate = c("AugTreePointer"),
  getMaxMapSizeEstimate = c("allowedSizeInMegs", "numRateCats"),
  checkAndCullMap = c("AugTreePointer", "allowedNumberOfElements", "cullProportion", "withinTransProbs",
                      "betweenTransProbs", "limProbs")
)

# routines that return nothing visible (void in C++)
.dmpcInvisible <- c("RestorePreviousConfig", "finalDeallocate", "checkAndCullMap")

local({
  ns <- topenv()
  for (entry in names(.dmpcEntryPoints)) {
    formalNames <- .dmpcEntryPoints[[entry]]
    callExpr <- as.call(c(list(as.name(".Call"), paste0("_DMphyClus_", entry)), lapply(formalNames, as.name),
                          list(PACKAGE = "DMphyClus")))
    body <- if (entry %in% .dmpcInvisible) call("invisible", callExpr) else callExpr
    stub <- as.function(c(setNames(rep(list(quote(expr = )), length(formalNames)), formalNames), list(body)), envir = ns)
    assign(entry, stub, envir = ns)
  }
})
