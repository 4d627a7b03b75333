#ifndef CVSHIM_HFPOSE2DARRAY_H
#define CVSHIM_HFPOSE2DARRAY_H
#include "../ros/ros.h"
#include "HFPose2D.h"
namespace handblobtracker {
struct HFPose2DArray {
    std_msgs::Header header;
    std::string id;
    std::vector<HFPose2D> measurements;
};
} // namespace handblobtracker
#endif
