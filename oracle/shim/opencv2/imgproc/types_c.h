/* oracle/shim/opencv2/imgproc/types_c.h -- TEST INFRASTRUCTURE ONLY: nothing of it is used by the solver sources. */
#pragma once
