// pose_math.hpp -- host-side restatement of the 3-vector pose arithmetic around the tracking hot path (plain doubles, no Eigen):
//   ComputeRelativePose / ComputeAbsolutePose            src/utils.cc:133-152
//   NormalizeAngle / RotationMatrix2D                    include/optimization_2d/normalize_angle.h:41-47, pose_graph_2d_error_term.h:44-51
//   Camera::ConvertCenterToPrincipal                     src/camera.cc:148-158
//   Camera::ConvertImagePlanePoseToCamera / ...ToRobot   src/camera.cc:160-175, :196-209, :224-231
//   MapBuilder::Tracking gate, UpdateCurrentPose, ComputeRelativeDA, keyframe test   src/map_builder.cc:42-57, :118-138, :157-166
// Used by nis_track_stream_keyframes (nis_api.cu); the GPU produces ComputePose's (pose, response), everything here is O(1) per frame.
#pragma once
#include <math.h>

#include "../../include/nislam.h"

namespace nis {
namespace pose {

struct P3 { double x, y, th; };

inline double normalize_angle(double a) {
  const double two_pi = 2.0 * M_PI;
  return a - two_pi * floor((a + M_PI) / two_pi);
}
inline P3 relative_pose(const P3& p1, const P3& p2) {        // Rw1^T (p2 - p1)
  const double c = cos(p1.th), s = sin(p1.th), dx = p2.x - p1.x, dy = p2.y - p1.y;
  return P3{c * dx + s * dy, -s * dx + c * dy, normalize_angle(p2.th - p1.th)};
}
inline P3 absolute_pose(const P3& p1, const P3& rel) {       // p1 + Rw1 rel
  const double c = cos(p1.th), s = sin(p1.th);
  return P3{p1.x + (c * rel.x - s * rel.y), p1.y + (s * rel.x + c * rel.y), normalize_angle(p1.th + rel.th)};
}
inline P3 center_to_principal(const nis_camera_model& cam, int W, int H, const P3& p) {   // p + (I - R(th)) O_bias
  const double c = cos(p.th), s = sin(p.th);
  const double ox = W * 0.5 - cam.cx, oy = H * 0.5 - cam.cy;
  return P3{p.x + ((1.0 - c) * ox + s * oy), p.y + (-s * ox + (1.0 - c) * oy), p.th};
}
inline P3 image_plane_to_camera(const nis_camera_model& cam, const P3& p) { return P3{p.x / cam.fx, p.y / cam.fy, p.th}; }
inline P3 camera_to_robot(const nis_camera_model& cam, const P3& p) {
  const double v[3] = {cam.height * p.x, cam.height * p.y, p.th};
  const double* E = cam.extrinsics;
  return P3{E[0] * v[0] + E[1] * v[1] + E[2] * v[2], E[3] * v[0] + E[4] * v[1] + E[5] * v[2], E[6] * v[0] + E[7] * v[1] + E[8] * v[2]};
}
inline P3 image_plane_to_robot(const nis_camera_model& cam, const P3& p) { return camera_to_robot(cam, image_plane_to_camera(cam, p)); }

// Camera::ConvertRobotPoseToImagePlane (src/camera.cc:211-222, :177-194, :233-241) and ConvertPrincipalToCenter (:136-146): the pose
// MapStitcher::AddImageToOccupancy places an image with (map_stitcher.cc:38-41)
inline void inverse3(const double* E, double* I) {       // cofactor inverse, like Eigen's fixed-size 3x3 inverse()
  const double c00 = E[4] * E[8] - E[5] * E[7], c01 = E[5] * E[6] - E[3] * E[8], c02 = E[3] * E[7] - E[4] * E[6];
  const double det = E[0] * c00 + E[1] * c01 + E[2] * c02, id = 1.0 / det;
  I[0] = c00 * id; I[1] = (E[2] * E[7] - E[1] * E[8]) * id; I[2] = (E[1] * E[5] - E[2] * E[4]) * id;
  I[3] = c01 * id; I[4] = (E[0] * E[8] - E[2] * E[6]) * id; I[5] = (E[2] * E[3] - E[0] * E[5]) * id;
  I[6] = c02 * id; I[7] = (E[1] * E[6] - E[0] * E[7]) * id; I[8] = (E[0] * E[4] - E[1] * E[3]) * id;
}
inline P3 robot_to_image_plane(const nis_camera_model& cam, const P3& r) {
  double I[9];
  inverse3(cam.extrinsics, I);
  P3 c{I[0] * r.x + I[1] * r.y + I[2] * r.th, I[3] * r.x + I[4] * r.y + I[5] * r.th, I[6] * r.x + I[7] * r.y + I[8] * r.th};
  c.x /= cam.height; c.y /= cam.height;
  return P3{cam.fx * c.x, cam.fy * c.y, c.th};
}
inline P3 principal_to_center(const nis_camera_model& cam, int W, int H, const P3& p) {   // p - (I - R(th)) O_bias
  const double c = cos(p.th), s = sin(p.th);
  const double ox = W * 0.5 - cam.cx, oy = H * 0.5 - cam.cy;
  return P3{p.x - ((1.0 - c) * ox + s * oy), p.y - (-s * ox + (1.0 - c) * oy), p.th};
}

// MapBuilder state that AddNewInput carries from frame to frame (include/map_builder.h)
struct TrackerState {
  bool init = false;
  P3 last_cf{0, 0, 0}, last_pose{0, 0, 0};        // of the last keyframe (UpdateIntermedium, map_builder.cc:99-106)
  P3 cur_cf{0, 0, 0}, cur_pose{0, 0, 0};
  double distance = 0.0;
};

// frame 0: MapBuilder::Initialize (map_builder.cc:86-97)
inline void initialize(const nis_camera_model& cam, TrackerState& st, nis_track_result& out) {
  st.cur_cf = P3{0, 0, 0};
  st.cur_pose = camera_to_robot(cam, image_plane_to_camera(cam, st.cur_cf));
  st.distance = 0.0;
  st.init = true;
  st.last_cf = st.cur_cf; st.last_pose = st.cur_pose;
  out.tracked = 1; out.inserted = 1; out.keyframe = -1; out.reserved = 0;
  for (int i = 0; i < 3; ++i) { out.response[i] = 0.0; out.relative_pose[i] = 0.0; }
}

// one AddNewInput step after ComputePose returned (pose_center, response); returns true when the frame is inserted as a keyframe
inline bool step(const nis_camera_model& cam, const nis_kfs_config& k, int W, int H, const double pose_center[3], const double response[3],
                 TrackerState& st, nis_track_result& out) {
  const P3 rel = center_to_principal(cam, W, H, P3{pose_center[0], pose_center[1], pose_center[2]});      // :131
  const bool good = response[0] > k.lower_response_thr && response[2] > k.lower_response_thr;             // :132
  bool inserted = false;
  if (good) {
    st.cur_cf = absolute_pose(st.last_cf, rel);                                                           // :134
    const P3 r0 = image_plane_to_robot(cam, st.last_cf), r1 = image_plane_to_robot(cam, st.cur_cf);       // :120-121
    st.cur_pose = absolute_pose(st.last_pose, relative_pose(r0, r1));                                     // :122-124
    const P3 d = image_plane_to_camera(cam, P3{st.cur_cf.x - st.last_cf.x, st.cur_cf.y - st.last_cf.y, st.cur_cf.th - st.last_cf.th});   // :159-160
    const double dist = sqrt(d.x * d.x + d.y * d.y), ang = fabs(d.th);                                    // :161-164
    const bool c1 = dist > k.max_distance, c2 = ang > k.max_angle;
    const bool c3 = response[0] > k.lower_response_thr && response[0] < k.upper_response_thr;
    const bool c4 = response[2] > k.lower_response_thr && response[2] < k.upper_response_thr;
    inserted = c1 || c2 || c3 || c4;                                                                      // :47-52
    if (inserted) {
      st.distance += dist;                                                                                // :53
      st.last_cf = st.cur_cf; st.last_pose = st.cur_pose;                                                 // UpdateIntermedium :68
    }
  }
  out.tracked = good ? 1 : 0; out.inserted = inserted ? 1 : 0; out.reserved = 0;
  out.relative_pose[0] = rel.x; out.relative_pose[1] = rel.y; out.relative_pose[2] = rel.th;
  for (int i = 0; i < 3; ++i) out.response[i] = response[i];
  return inserted;
}

inline void snapshot(const TrackerState& st, nis_track_result& out) {
  out.cf_pose[0] = st.cur_cf.x; out.cf_pose[1] = st.cur_cf.y; out.cf_pose[2] = st.cur_cf.th;
  out.pose[0] = st.cur_pose.x; out.pose[1] = st.cur_pose.y; out.pose[2] = st.cur_pose.th;
  out.distance = st.distance;
}

}  // namespace pose
}  // namespace nis
