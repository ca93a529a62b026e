// Host-built "plan" that maps the reference's four wave types (Rc, Rg, Lc, Lg;
// /root/reference/src/SWD/surfdisp.cpp:190-297) onto unique period sequences and eigen solves.
// Shared by the two translation units (capi.cu, swd_roots_tu.cu).
#pragma once
#include "swd_roots.cuh"

namespace rfs {

#define RFS_MAX_SEQ 12
#define RFS_MAX_ROWS 4

// one requested data block ("row"): a wave type on a period list
struct SwdRow {
  int type;      // 0 Rc, 1 Rg, 2 Lc, 3 Lg
  int nper;      // periods
  int per_off;   // offset of its period list in the period table
  int s0, s1, s2;  // sequences: T, 1.05 T, 0.95 T (s1,s2 = -1 for phase velocity)
  int d_off;     // offset of this row in the concatenated data vector
};

struct SwdPlan {
  int nseq, nrow;
  int nsolve;  // total (sequence, period) pairs == total periods over sequences
  int ndata;   // total data count over rows
  int nmode;   // modes solved (mode+1); the last one is reported
  SwdSeq seq[RFS_MAX_SEQ];
  SwdRow row[RFS_MAX_ROWS];
};

}  // namespace rfs
