// Kernel instantiations, group 1 (split across translation units so they compile in parallel).
#define DKG_GROUP 1
#define DKG_GROUP_SHAPES(X) X(12,11) X(6,3) X(12,2) X(16,2)
#define DKG_GROUP_GROUPED_SHAPES(X) X(12,11) X(6,3) X(12,2) X(16,2)
#define DKG_GROUP_NSQ_SHAPES(X) X(6,3) X(12,2) X(16,2) X(12,6)
#define DKG_GROUP_NSQ_BG_SHAPES(X) X(16,6)
#include "dkg_kernels.inc"
