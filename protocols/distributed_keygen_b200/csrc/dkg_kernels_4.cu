// Kernel instantiations, group 4 (split across translation units so they compile in parallel).
#define DKG_GROUP 4
#define DKG_GROUP_SHAPES(X) X(16,17) X(22,3) X(16,6)
#define DKG_GROUP_GROUPED_SHAPES(X) X(22,3) X(16,6)
#define DKG_GROUP_NSQ_SHAPES(X) X(22,3) X(16,6)
#define DKG_GROUP_NSQ_BG_SHAPES(X) X(16,9)
#include "dkg_kernels.inc"
