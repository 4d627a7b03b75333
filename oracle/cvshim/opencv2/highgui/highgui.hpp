// cvshim: nothing of highgui is used by the hot-path sources
#include "../core/core.hpp"
