#ifndef CVSHIM_ROIARRAY_H
#define CVSHIM_ROIARRAY_H
#include "../ros/ros.h"
#include "../sensor_msgs/RegionOfInterest.h"
namespace facetracking {
struct ROIArray {
    std_msgs::Header header;
    std::vector<sensor_msgs::RegionOfInterest> ROIs;
    std::vector<int> ids;
};
typedef std::shared_ptr<ROIArray> ROIArrayPtr;
typedef std::shared_ptr<ROIArray const> ROIArrayConstPtr;
} // namespace facetracking
#endif
