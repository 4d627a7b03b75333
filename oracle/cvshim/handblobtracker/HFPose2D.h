#ifndef CVSHIM_HFPOSE2D_H
#define CVSHIM_HFPOSE2D_H
namespace handblobtracker {
struct HFPose2D {
    double x, y;
    HFPose2D() : x(0), y(0) {}
};
} // namespace handblobtracker
#endif
