// Kernel instantiations, group 2 (split across translation units so they compile in parallel).
#define DKG_GROUP 2
#define DKG_GROUP_SHAPES(X) X(16,16) X(12,3) X(16,3)
#define DKG_GROUP_GROUPED_SHAPES(X) X(12,3) X(16,3) X(14,5)
#define DKG_GROUP_NSQ_SHAPES(X) X(12,3) X(16,3) X(14,5)
#include "dkg_kernels.inc"
