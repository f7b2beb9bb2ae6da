// Stand-in for <pcl/filters/extract_indices.h> (not installed here): everything the ltremovert sources use lives in ltr_shim_core.h.
#include "ltr_shim_core.h"
