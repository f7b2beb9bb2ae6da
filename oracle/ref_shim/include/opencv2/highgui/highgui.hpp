// Stand-in for <opencv2/highgui/highgui.hpp> (not installed here): everything the ltremovert sources use lives in ltr_shim_core.h.
#include "ltr_shim_core.h"
