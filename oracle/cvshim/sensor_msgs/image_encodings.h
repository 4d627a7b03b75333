#ifndef CVSHIM_IMAGE_ENCODINGS_H
#define CVSHIM_IMAGE_ENCODINGS_H
#include <string>
namespace sensor_msgs {
namespace image_encodings {
const std::string RGB8 = "rgb8";
const std::string MONO8 = "mono8";
} // namespace image_encodings
} // namespace sensor_msgs
#endif
