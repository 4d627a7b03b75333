#ifndef CVSHIM_CV_BRIDGE_H
#define CVSHIM_CV_BRIDGE_H
#include "../opencv2/core/core.hpp"
#include "../sensor_msgs/Image.h"
namespace cv_bridge {
class CvImage {
  public:
    std_msgs::Header header;
    std::string encoding;
    cv::Mat image;
    sensor_msgs::ImagePtr toImageMsg() const
    {
        sensor_msgs::ImagePtr m(new sensor_msgs::Image);
        m->header = header;
        m->encoding = encoding;
        m->height = image.rows;
        m->width = image.cols;
        m->step = (uint32_t)(image.elemSize() * image.cols);
        m->data.resize((size_t)m->step * m->height);
        for (int r = 0; r < image.rows; r++)
            std::memcpy(&m->data[(size_t)r * m->step], image.data + (size_t)r * image.step, m->step);
        return m;
    }
};
typedef std::shared_ptr<CvImage> CvImagePtr;
inline CvImagePtr toCvCopy(const sensor_msgs::ImageConstPtr& msg, const std::string& encoding)
{
    CvImagePtr p(new CvImage);
    p->header = msg->header;
    p->encoding = encoding;
    const int cn = encoding == "rgb8" ? 3 : 1;
    p->image.create((int)msg->height, (int)msg->width, CV_MAKETYPE(CV_8U, cn));
    for (int r = 0; r < p->image.rows; r++)
        std::memcpy(p->image.data + (size_t)r * p->image.step, &msg->data[(size_t)r * msg->step],
                    (size_t)msg->width * cn);
    return p;
}
} // namespace cv_bridge
#endif
