#ifndef CVSHIM_ROI_H
#define CVSHIM_ROI_H
#include <cstdint>
namespace sensor_msgs {
struct RegionOfInterest {
    uint32_t x_offset, y_offset, height, width;
    uint8_t do_rectify;
    RegionOfInterest() : x_offset(0), y_offset(0), height(0), width(0), do_rectify(0) {}
};
} // namespace sensor_msgs
#endif
