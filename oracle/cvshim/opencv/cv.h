// cvshim: legacy umbrella header
#include "../opencv2/core/core.hpp"
