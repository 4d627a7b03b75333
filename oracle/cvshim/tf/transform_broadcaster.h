#ifndef CVSHIM_TF_BROADCASTER_H
#define CVSHIM_TF_BROADCASTER_H
#include "../ros/ros.h"
namespace tf {
struct Vector3 {
    double x, y, z;
    Vector3() : x(0), y(0), z(0) {}
    Vector3(double a, double b, double c) : x(a), y(b), z(c) {}
};
struct Quaternion {
    double yaw, pitch, roll; // kept as the Euler triple handed to setEuler
    Quaternion() : yaw(0), pitch(0), roll(0) {}
    void setEuler(double y, double p, double r)
    {
        yaw = y;
        pitch = p;
        roll = r;
    }
};
struct Transform {
    Vector3 origin;
    Quaternion rotation;
    void setOrigin(const Vector3& v) { origin = v; }
    void setRotation(const Quaternion& q) { rotation = q; }
};
struct StampedTransform : Transform {
    ros::Time stamp;
    std::string frame_id, child_frame_id;
    StampedTransform(const Transform& t, const ros::Time& s, const std::string& f, const std::string& c)
        : Transform(t), stamp(s), frame_id(f), child_frame_id(c)
    {
    }
};
namespace shim {
inline std::vector<StampedTransform>& sent()
{
    static std::vector<StampedTransform> v;
    return v;
}
} // namespace shim
class TransformBroadcaster {
  public:
    void sendTransform(const StampedTransform& t) { shim::sent().push_back(t); }
};
} // namespace tf
#endif
