// cvshim/ros/ros.h -- TEST INFRASTRUCTURE: the sliver of roscpp that src/pfPose.{h,cpp} touches, with every
// publication recorded instead of sent, so PFTracker can be driven without a ROS master (see ref_glue.cpp).
#ifndef CVSHIM_ROS_H
#define CVSHIM_ROS_H
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace boost {
using std::shared_ptr;
template <class... A>
int bind(A...)
{
    return 0;
}
} // namespace boost
static const int _1 = 1, _2 = 2, _3 = 3; // boost::bind placeholders

namespace ros {
struct Time {
    double sec;
    Time() : sec(0) {}
};
namespace shim {
inline std::map<std::string, std::string>& params()
{
    static std::map<std::string, std::string> p;
    return p;
}
inline std::string& package_path()
{
    static std::string p;
    return p;
}
template <class M>
std::vector<M>& published() // every message of type M published through ros::Publisher, in order
{
    static std::vector<M> v;
    return v;
}
} // namespace shim
class Publisher {
  public:
    template <class M>
    void publish(const M& m) const
    {
        shim::published<M>().push_back(m);
    }
};
class NodeHandle {
  public:
    template <class M>
    Publisher advertise(const std::string&, int)
    {
        return Publisher();
    }
};
namespace param {
template <class T>
bool param(const std::string& name, T& val, const T& def)
{
    auto it = shim::params().find(name);
    val = (it == shim::params().end()) ? def : T(it->second);
    return it != shim::params().end();
}
} // namespace param
inline void init(int&, char**, const std::string&) {}
inline void spin() {}
} // namespace ros
#define ROS_INFO(...) \
    do {              \
    } while (0)

namespace std_msgs {
struct Header {
    uint32_t seq;
    ros::Time stamp;
    std::string frame_id;
    Header() : seq(0) {}
};
} // namespace std_msgs
#endif
