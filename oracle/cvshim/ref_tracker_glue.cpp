// oracle/cvshim/ref_tracker_glue.cpp -- TEST INFRASTRUCTURE: drives the reference's own PFTracker
// (src/pfPose.{h,cpp}, compiled in place) without ROS: synthetic image / likelihood / face-ROI messages go in,
// everything it publishes (2-D joints, TF tree, probability image) and its filter states come out.
//
// Two builds (oracle/Makefile):
//   _ref/libref.so         pfPose.cpp on the reference's own KF_model / my_gmm / pf2DRao            (CPU)
//   _ref/libref_dropin.so  the same unmodified pfPose.cpp on include/mkf_shims.hpp + libmkf_b200.so  (GPU), built with
//                          -DMKF_DROPIN -D__PF2DRAO -D__MYGMM -D__KFMODEL (the reference headers' own include guards,
//                          so "pf2DRao.h" contributes nothing) and -include mkf_shims.hpp: INTEGRATION.md section 2
//                          carried out for real.
// everything pfPose.h includes is pulled in first (include guards), so that the access override below only
// touches the PFTracker class body and not the standard library
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>

#include "opencv2/highgui/highgui.hpp"
#include "opencv2/objdetect/objdetect.hpp"
#include "opencv2/ml/ml.hpp"
#include <opencv/cv.h>
#include <ros/ros.h>
#include <cv_bridge/cv_bridge.h>
#include "opencv2/video/tracking.hpp"
#include <opencv2/features2d/features2d.hpp>
#include <image_transport/image_transport.h>
#include <sensor_msgs/image_encodings.h>
#include "sensor_msgs/Image.h"
#include "sensor_msgs/RegionOfInterest.h"
#include "geometry_msgs/Point.h"
#include <message_filters/subscriber.h>
#include <message_filters/time_synchronizer.h>
#include "pf2DRao.h"
#include <ros/package.h>
#include "handblobtracker/HFPose2D.h"
#include "handblobtracker/HFPose2DArray.h"
#include "facetracking/ROIArray.h"
#include <tf/transform_broadcaster.h>

#define private public
#include "pfPose.h"
#undef private

namespace cv {
void cvshim_push_tick(int64 t);
}

extern "C" {

// PFTracker::PFTracker (src/pfPose.cpp:7-74): models <pkg_path><left_file> / <pkg_path><right_file>; the two ticks
// are what cv::getTickCount() returns inside the constructor's two resample() calls
void* ref_tracker_create(const char* pkg_path, const char* left_file, const char* right_file, int64_t tick1, int64_t tick2)
{
    ros::shim::package_path() = pkg_path;
    ros::shim::params()["left_arm_training"] = left_file;
    ros::shim::params()["right_arm_training"] = right_file;
    cv::cvshim_push_tick(tick1);
    cv::cvshim_push_tick(tick2);
    return new PFTracker();
}
void ref_tracker_destroy(void* tr) { delete (PFTracker*)tr; }
#ifdef MKF_DROPIN
// alias_mode for the ParticleFilter objects PFTracker constructs from now on; candidates drawn by the shim's
// getSamples() are appended to the same log as the CPU build's cv::randn draws (interleaved x, y)
void ref_dropin_configure(int alias_mode)
{
    mkf::default_params().alias_mode = alias_mode;
    mkf::sample_hook() = [](const cv::Mat& m) {
        std::vector<double> v((size_t)2 * m.cols);
        for (int i = 0; i < m.cols; i++) {
            v[2 * i] = m.at<double>(0, i);
            v[2 * i + 1] = m.at<double>(1, i);
        }
        cv::cvshim_random_log().push_back(v);
    };
}
#endif
int ref_tracker_num_particles(void* tr) { return ((PFTracker*)tr)->numParticles; }

void ref_set_rng_seed(uint64_t seed) { cv::cvshim_seed_the_rng(seed); }
void ref_random_log_clear() { cv::cvshim_random_log().clear(); }
int ref_random_log_count() { return (int)cv::cvshim_random_log().size(); }
int ref_random_log_get(int i, double* out, int cap)
{
    const std::vector<double>& v = cv::cvshim_random_log()[i];
    const int n = (int)v.size() < cap ? (int)v.size() : cap;
    std::memcpy(out, v.data(), sizeof(double) * n);
    return (int)v.size();
}

// PFTracker::callback (src/pfPose.cpp:332-385) on one synthetic frame.  like: rows x cols uint8 (not blurred);
// roi: x_offset, y_offset, height, width when has_roi; ticks: the cv::getTickCount() values seen by the six
// resample() calls of the frame (candidates arm 1, arm 2, then indicator + posterior for each arm).
void ref_tracker_callback(void* trv, const uint8_t* like, int rows, int cols, int has_roi, const uint32_t* roi,
                          const int64_t* ticks, int nticks)
{
    PFTracker* tr = (PFTracker*)trv;
    sensor_msgs::ImagePtr im(new sensor_msgs::Image), lk(new sensor_msgs::Image);
    im->height = lk->height = rows;
    im->width = lk->width = cols;
    im->encoding = "rgb8";
    im->step = cols * 3;
    im->data.assign((size_t)rows * cols * 3, 0);
    lk->encoding = "mono8";
    lk->step = cols;
    lk->data.assign(like, like + (size_t)rows * cols);
    facetracking::ROIArrayPtr rm(new facetracking::ROIArray);
    if (has_roi) {
        sensor_msgs::RegionOfInterest r;
        r.x_offset = roi[0];
        r.y_offset = roi[1];
        r.height = roi[2];
        r.width = roi[3];
        rm->ROIs.push_back(r);
    }
    for (int i = 0; i < nticks; i++) cv::cvshim_push_tick(ticks[i]);
    tf::shim::sent().clear();
    ros::shim::published<handblobtracker::HFPose2DArray>().clear();
    tr->callback(im, lk, rm);
}

// what the last callback published: joints2d 8 x 2 (publish2Dpos), tf 10 x 3 (nine translations in broadcast order
// + the camera Euler triple), prob: rows x cols blurred likelihood (channel 0 of /probImage).  returns #transforms.
int ref_tracker_outputs(double* joints2d, double* tf10x3, uint8_t* prob, int rows, int cols)
{
    const std::vector<handblobtracker::HFPose2DArray>& pub = ros::shim::published<handblobtracker::HFPose2DArray>();
    if (joints2d && !pub.empty())
        for (size_t i = 0; i < pub.back().measurements.size() && i < 8; i++) {
            joints2d[2 * i] = pub.back().measurements[i].x;
            joints2d[2 * i + 1] = pub.back().measurements[i].y;
        }
    const std::vector<tf::StampedTransform>& s = tf::shim::sent();
    if (tf10x3)
        for (size_t i = 0; i < s.size() && i < 9; i++) {
            tf10x3[3 * i] = s[i].origin.x;
            tf10x3[3 * i + 1] = s[i].origin.y;
            tf10x3[3 * i + 2] = s[i].origin.z;
            if (i == 8) {
                tf10x3[27] = s[i].rotation.yaw;
                tf10x3[28] = s[i].rotation.pitch;
                tf10x3[29] = s[i].rotation.roll;
            }
        }
    if (prob) {
        auto it = image_transport::shim::last_image().find("/probImage");
        if (it != image_transport::shim::last_image().end() && it->second)
            for (int r = 0; r < rows; r++)
                for (int c = 0; c < cols; c++) prob[(size_t)r * cols + c] = it->second->data[(size_t)r * it->second->step + 3 * c];
    }
    return (int)s.size();
}

void ref_tracker_get_state(void* trv, int arm, double* x, double* P)
{
    PFTracker* tr = (PFTracker*)trv;
    ParticleFilter* pf = arm ? tr->pf2 : tr->pf1;
    const int N = pf->gmm.nParticles;
#ifdef MKF_DROPIN
    pf->gmm.syncTracks(); // the state lives on the device
#endif
    for (int j = 0; j < N; j++) {
        const cv::Mat& s = pf->gmm.tracks[j].state;
        const cv::Mat& c = pf->gmm.tracks[j].cov;
        const int d = s.rows;
        if (x)
            for (int i = 0; i < d; i++) x[(size_t)j * d + i] = s.at<double>(i, 0);
        if (P)
            for (int r = 0; r < d; r++)
                for (int q = 0; q < d; q++) P[((size_t)j * d + r) * d + q] = c.at<double>(r, q);
    }
}

// e = h_pca.t()*getEstimator() + m_pca.t() for one arm (src/pfPose.cpp:347-348) and PFTracker::get3Dpose of it
void ref_tracker_pose(void* trv, int arm, double* e22, double* pos3d)
{
    PFTracker* tr = (PFTracker*)trv;
    cv::Mat e = arm ? cv::Mat(tr->h2_pca.t() * tr->pf2->getEstimator() + tr->m2_pca.t())
                    : cv::Mat(tr->h1_pca.t() * tr->pf1->getEstimator() + tr->m1_pca.t());
    if (e22)
        for (int i = 0; i < e.rows; i++) e22[i] = e.at<double>(i, 0);
    if (pos3d) {
        cv::Mat p = tr->get3Dpose(e);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 5; c++) pos3d[r * 5 + c] = p.at<double>(r, c);
    }
}

} // extern "C"
