#ifndef CVSHIM_IMAGE_TRANSPORT_H
#define CVSHIM_IMAGE_TRANSPORT_H
#include "../sensor_msgs/Image.h"
namespace image_transport {
namespace shim {
inline std::map<std::string, sensor_msgs::ImagePtr>& last_image() // last image published per topic
{
    static std::map<std::string, sensor_msgs::ImagePtr> m;
    return m;
}
} // namespace shim
class Publisher {
  public:
    std::string topic;
    void publish(const sensor_msgs::ImagePtr& m) const { shim::last_image()[topic] = m; }
};
class ImageTransport {
  public:
    explicit ImageTransport(ros::NodeHandle&) {}
    Publisher advertise(const std::string& topic, int)
    {
        Publisher p;
        p.topic = topic;
        return p;
    }
};
} // namespace image_transport
#endif
