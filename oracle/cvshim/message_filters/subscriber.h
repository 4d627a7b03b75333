#ifndef CVSHIM_MF_SUBSCRIBER_H
#define CVSHIM_MF_SUBSCRIBER_H
#include "../ros/ros.h"
namespace message_filters {
template <class M>
class Subscriber {
  public:
    void subscribe(ros::NodeHandle&, const std::string&, int) {}
};
} // namespace message_filters
#endif
