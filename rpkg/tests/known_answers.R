## Known answers for a box with R: run after `R CMD INSTALL` of the assembled package (INTEGRATION.md), e.g.
##   Rscript rpkg/tests/known_answers.R
## The six log-likelihoods below are those of the reference's own translation units on the packaged data
## (tests/golden/ref_vectors.npz "vignette/logliks", produced by tests/golden/make_ref_vectors.py; equal to the
## independently derived values of BASELINE.md section 5): the vignette alignment (data/seqsToCluster.rda, 6 x 1617),
## the tree (((1,2),3),((4,5),6)) in alignment-row order, limProbs = (a .39, t .22, c .17, g .22), the packaged
## transition matrices.  The call is the upstream R function unchanged (R/DMphyClusInterface.R:216-252), so this
## checks the whole drop-in path: getConvertedAlignment -> logLikCpp -> finalDeallocate through rpkg/src and libdmpc.
## NOT run in this repository's build image (no R there); tests/test_oracle_golden.py::test_oracle_vignette_known_answers
## and tests/test_gpu_parity.py check the same six values through the C-ABI.
library(DMphyClus)
data("seqsToCluster", package = "DMphyClus")
data("withinClusTransProbs", package = "DMphyClus")
data("betweenClusTransProbs", package = "DMphyClus")
seqs <- as.character(ape::as.matrix.DNAbin(seqsToCluster))
stopifnot(all(dim(seqs) == c(6, 1617)))

edge <- matrix(c(7, 8,  8, 9,  9, 1,  9, 2,  8, 3,  7, 10,  10, 11,  11, 4,  11, 5,  10, 6), ncol = 2, byrow = TRUE)
storage.mode(edge) <- "integer"
phylo <- structure(list(edge = edge, tip.label = rownames(seqs), Nnode = 5L, edge.length = rep(1, nrow(edge))),
                   class = "phylo")
limProbs <- c(a = 0.39, t = 0.22, c = 0.17, g = 0.22)
named <- function(x) { names(x) <- rownames(seqs); x }

cases <- list(
  list(clus = named(c(1, 1, 1, 2, 2, 2)), w = 5,  b = 5,  want = -3072.967109235698),
  list(clus = named(c(1, 1, 1, 2, 2, 2)), w = 1,  b = 1,  want = -3148.266004654241),
  list(clus = named(c(1, 1, 1, 2, 2, 2)), w = 10, b = 10, want = -3038.497338364486),
  list(clus = named(c(1, 1, 1, 2, 2, 2)), w = 5,  b = 1,  want = -3130.306408470979),
  list(clus = named(rep(1, 6)),           w = 5,  b = 5,  want = -3510.074064616059),   # one cluster: all within
  list(clus = named(1:6),                 w = 5,  b = 5,  want = -3259.751644946885))   # six singletons: all between

for (k in seq_along(cases)) {
  cs <- cases[[k]]
  got <- logLikFromClusInd(phylogeny = phylo, betweenTransMatList = betweenClusTransProbs[[cs$b]],
                           withinTransMatList = withinClusTransProbs[[cs$w]], clusInd = cs$clus, limProbs = limProbs,
                           numLikThreads = 1, alignment = seqs, basicChecks = FALSE)
  cat(sprintf("case %d: %.12f (expected %.12f)\n", k, got, cs$want))
  stopifnot(abs(got - cs$want) <= 1e-9 * abs(cs$want))
}
cat("known answers: all six agree to 1e-9 relative\n")
