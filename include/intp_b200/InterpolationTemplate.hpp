// intp_b200/InterpolationTemplate.hpp -- the reference keeps InterpolationFunctionTemplate in its own
// header (src/include/InterpolationTemplate.hpp:32-604) and includes it from Interpolation.hpp; here
// both classes live in Interpolation.hpp and this file only keeps the include name working.
#ifndef INTP_B200_INTERPOLATION_TEMPLATE_HPP
#define INTP_B200_INTERPOLATION_TEMPLATE_HPP
#include "Interpolation.hpp"
#endif
