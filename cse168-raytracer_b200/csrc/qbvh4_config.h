// Build-time switches of the QBVH4 node encoding and its decode -- shared by the builders (bvh_build.cpp, lbvh_impl.cuh) and the
// traversal (traverse.cuh), which must agree.  The defaults are the product; the alternatives exist for A/B measurements
// (MIROGPU_NVCC_DEFS="-DMIRO_QDIRECT=0" python -m ... _build, see tools/gpu_ab.sh).
#ifndef MIROGPU_QBVH4_CONFIG_H
#define MIROGPU_QBVH4_CONFIG_H

// 0 (default): plane bytes become subnormal binary16 values q 2^-24, converted to binary32 (one PRMT per pair on the ALU pipe + one
//    conversion each on the FMA pipe); the node stores origin and 2^24 cell.
// 1: plane bytes become binary32 values 1 + q 2^-15 with one PRMT each and no conversion; the node stores origin - 2^15 cell and
//    2^15 cell.  12 instructions fewer per node step (143 -> 132) and 3 % SLOWER on the bench step (8.97 -> 8.69 Grays/s): the
//    kernel is bound by the ALU pipe (PRMT, min / max, selects, compares: ~70 of a node step's ~145 instructions at half rate), and
//    this trades 24 FMA-pipe conversions for 12 more PRMTs.  Kept selectable; hits are identical either way (same tests).
#ifndef MIRO_QDIRECT
#define MIRO_QDIRECT 0
#endif
#define MIRO_QSHIFT (MIRO_QDIRECT ? 15 : 24)

// 1: plane distances of two children per packed binary32 FMA (FFMA2).
#ifndef MIRO_FFMA2
#define MIRO_FFMA2 1
#endif
// 1: the cell sizes are read ready-made from the node's last two words; 0: formed from the exponent bytes (needs MIRO_QDIRECT 0).
#ifndef MIRO_QCELL
#define MIRO_QCELL 1
#endif
#if MIRO_QDIRECT && !MIRO_QCELL
#error "MIRO_QDIRECT needs MIRO_QCELL"
#endif

#endif
