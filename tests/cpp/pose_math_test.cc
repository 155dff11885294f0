// Drives ni_slam_b200/host/pose_math.hpp (the host half of nis_track_stream_keyframes) with scripted ComputePose outputs so the
// CPU suite can compare it with the Python restatement of map_builder.cc:30-70 (oracle/tracker_ref.py) -- no GPU involved.
// stdin:  fx fy cx cy height E[9]  max_distance max_angle lower upper  W H  n   then n lines: r0 r1 r2 px py pth
// stdout: one line per frame (frame 0 = Initialize): tracked inserted keyframe cf[3] pose[3] distance rel[3]
#include <stdio.h>

#include "../../ni_slam_b200/host/pose_math.hpp"

int main() {
  nis_camera_model cam;
  nis_kfs_config k;
  int W, H, n;
  if (scanf("%lf %lf %lf %lf %lf", &cam.fx, &cam.fy, &cam.cx, &cam.cy, &cam.height) != 5) return 1;
  for (int i = 0; i < 9; ++i) if (scanf("%lf", &cam.extrinsics[i]) != 1) return 1;
  if (scanf("%lf %lf %lf %lf %d %d %d", &k.max_distance, &k.max_angle, &k.lower_response_thr, &k.upper_response_thr, &W, &H, &n) != 7) return 1;
  nis::pose::TrackerState st;
  nis_track_result o;
  nis::pose::initialize(cam, st, o);
  nis::pose::snapshot(st, o);
  int K = 0;
  printf("%d %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g 0 0 0\n", o.tracked, o.inserted, -1, o.cf_pose[0], o.cf_pose[1], o.cf_pose[2],
         o.pose[0], o.pose[1], o.pose[2], o.distance);
  for (int i = 1; i <= n; ++i) {
    double r[3], p[3];
    if (scanf("%lf %lf %lf %lf %lf %lf", &r[0], &r[1], &r[2], &p[0], &p[1], &p[2]) != 6) return 1;
    const int kf = K;
    if (nis::pose::step(cam, k, W, H, p, r, st, o)) K = i;
    nis::pose::snapshot(st, o);
    printf("%d %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", o.tracked, o.inserted, kf, o.cf_pose[0], o.cf_pose[1],
           o.cf_pose[2], o.pose[0], o.pose[1], o.pose[2], o.distance, o.relative_pose[0], o.relative_pose[1], o.relative_pose[2]);
  }
  return 0;
}
