// Kernel instantiations, group 5 (split across translation units so they compile in parallel).
#define DKG_GROUP 5
#define DKG_GROUP_SHAPES(X) X(16,9) X(16,12)
#define DKG_GROUP_NSQ_SHAPES(X) X(16,9) X(12,11) X(14,7)
#define DKG_GROUP_NSQ_BG_SHAPES(X) X(14,7)
#include "dkg_kernels.inc"
