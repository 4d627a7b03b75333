#ifndef CVSHIM_GEOMETRY_POINT_H
#define CVSHIM_GEOMETRY_POINT_H
namespace geometry_msgs {
struct Point {
    double x, y, z;
};
} // namespace geometry_msgs
#endif
