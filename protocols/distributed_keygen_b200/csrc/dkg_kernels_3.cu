// Kernel instantiations, group 3 (split across translation units so they compile in parallel).
#define DKG_GROUP 3
#define DKG_GROUP_SHAPES(X) X(20,13) X(16,4) X(16,5)
#define DKG_GROUP_GROUPED_SHAPES(X) X(16,4) X(16,5)
#define DKG_GROUP_NSQ_SHAPES(X) X(16,4) X(16,5) X(18,4)
#define DKG_GROUP_NSQ_BG_SHAPES(X) X(12,11)
#include "dkg_kernels.inc"
