// The reference installs its headers as <BSplineInterpolation/util.hpp> (CMakeLists.txt:38-40); this
// forwarder keeps that include line working with -I<repo>/include.
#include "../intp_b200/util.hpp"
