// gsl/gsl_sf_gamma.h stand-in (included by IntermediateNode.cpp:1, nothing from it is used)
