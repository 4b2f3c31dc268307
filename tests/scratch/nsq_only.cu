#define DKG_GROUP 9
#define DKG_GROUP_SHAPES(X)
#define DKG_GROUP_NSQ_SHAPES(X) X(14,5)
#include "../../protocols/distributed_keygen_b200/csrc/dkg_kernels.inc"
