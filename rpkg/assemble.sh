#!/bin/sh
# Assembles an installable DMphyClus source package whose likelihood runs in libdmpc:
#   the upstream package's R code, data, man pages and vignette — UNCHANGED — plus this directory's DESCRIPTION,
#   NAMESPACE, configure, R/RcppExports.R and src/ (the upstream src/ is dropped: AugTree.*, IntermediateNode.*,
#   InputNode.*, TreeNode.h, helper.h and its Makevars are no longer compiled).
#
#   sh rpkg/assemble.sh /path/to/upstream/DMphyClus /tmp/DMphyClus.b200
#   DMPC_HOME=$PWD R CMD INSTALL /tmp/DMphyClus.b200
#   DMPC_DEVICES=0,1,2,3,4,5,6,7 Rscript -e 'rmarkdown::render("/tmp/DMphyClus.b200/vignettes/DMphyClus_Example.Rmd")'
#
# (Optional, on a box with R: Rscript -e 'Rcpp::compileAttributes("/tmp/DMphyClus.b200")' regenerates
#  src/RcppExports.cpp and R/RcppExports.R from the [[Rcpp::export]] attributes; the hand-written ones are equivalent.)
set -e
up=${1:?usage: assemble.sh <upstream package dir> <output dir>}
out=${2:?usage: assemble.sh <upstream package dir> <output dir>}
here=$(cd "$(dirname "$0")" && pwd)
for d in R data man vignettes; do
  [ -d "$up/$d" ] || { echo "assemble: $up/$d is missing (not a DMphyClus source tree?)" >&2; exit 1; }
done
rm -rf "$out"
mkdir -p "$out/src" "$out/R"
cp -R "$up/data" "$up/man" "$up/vignettes" "$out/"
for f in "$up"/R/*.R; do
  case "$(basename "$f")" in RcppExports.R) ;; *) cp "$f" "$out/R/";; esac
done
[ -f "$up/README.md" ] && cp "$up/README.md" "$out/README.upstream.md"
cp "$here/DESCRIPTION" "$here/NAMESPACE" "$here/configure" "$out/"
cp "$here/R/RcppExports.R" "$out/R/"
cp "$here/src/clusterFunArmadillo.cpp" "$here/src/RcppExports.cpp" "$here/src/Makevars.in" "$here/src/Makevars" "$out/src/"
chmod +x "$out/configure"
echo "assembled $out (R code, data, man, vignettes from $up; compiled code from $here)"
