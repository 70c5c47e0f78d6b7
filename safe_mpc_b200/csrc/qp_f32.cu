// The QP solver of qp.cu in its fp32-storage flavour (namespace smpc::f32): everything a solve streams -- stage records, search
// directions, condensed matrices, Riccati factors -- is stored in fp32, the iterate and all arithmetic stay fp64 (qp_split.cuh).
// Selected per handle by smpc_problem_t::precision = SMPC_PREC_F32.
#define QS_REAL float
#define QS_FLAVOUR f32
#define QS_OTHER_FLAVOUR f64
#include "qp.cu"
