// Build configuration for compiling the reference's vendored Ceres 2.0.0
// (/root/reference/thirdparty/ceres-solver) with oracle/ref/Makefile instead
// of its CMake build.  Mirrors the options SURVEY.md §8(c) measured with:
// MINIGLOG, EIGENSPARSE, C++ threads, no SuiteSparse/CXSparse/LAPACK.
// This file is test infrastructure (oracle/), never part of the product.
#ifndef CERES_PUBLIC_INTERNAL_CONFIG_H_
#define CERES_PUBLIC_INTERNAL_CONFIG_H_
#define CERES_USE_EIGEN_SPARSE
#define CERES_NO_LAPACK
#define CERES_NO_SUITESPARSE
#define CERES_NO_CXSPARSE
#define CERES_NO_ACCELERATE_SPARSE
#define CERES_RESTRICT_SCHUR_SPECIALIZATION
#define CERES_USE_CXX_THREADS
#endif
