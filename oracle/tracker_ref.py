"""CPU restatement of MapBuilder::AddNewInput's tracking / keyframe policy -- TEST INFRASTRUCTURE ONLY (only tests/, smoke() and
bench.py's cpu_baseline may import this).  Parity unpinned by the reference (it ships no tests); pinned by the known answers in
tests/test_oracle.py.

Follows, line by line:
  MapBuilder::AddNewInput / Initialize / UpdateIntermedium / UpdateCurrentPose / Tracking / ComputeRelativeDA
                                                     src/map_builder.cc:30-70, :86-106, :118-138, :157-166
  ComputeRelativePose / ComputeAbsolutePose          src/utils.cc:133-152
  NormalizeAngle / RotationMatrix2D                  include/optimization_2d/normalize_angle.h:41-47, pose_graph_2d_error_term.h:44-51
  Camera::ConvertCenterToPrincipal / ConvertImagePlanePoseToCamera / ConvertCameraPoseToRobot / ConvertImagePlanePoseToRobot
                                                     src/camera.cc:148-158, :160-175, :196-209, :224-231
Loop closure, optimisation and stitching (map_builder.cc:59-66) are outside the path and left out.
"""
import math

import numpy as np


def rotation_matrix_2d(yaw):
    c, s = math.cos(yaw), math.sin(yaw)
    return np.array([[c, -s], [s, c]], np.float64)


def normalize_angle(a):
    two_pi = 2.0 * math.pi
    return a - two_pi * math.floor((a + math.pi) / two_pi)


def compute_relative_pose(p1, p2):                      # utils.cc:133-141
    r = np.zeros(3)
    r[:2] = rotation_matrix_2d(p1[2]).T @ (p2[:2] - p1[:2])
    r[2] = normalize_angle(p2[2] - p1[2])
    return r


def compute_absolute_pose(p1, rel):                     # utils.cc:143-152
    r = np.zeros(3)
    r[:2] = p1[:2] + rotation_matrix_2d(p1[2]) @ rel[:2]
    r[2] = normalize_angle(p1[2] + rel[2])
    return r


class Camera:
    """The members the pose conversions read: _new_K, _image_width/_image_height, _height, _extrinsics."""

    def __init__(self, fx, fy, cx, cy, height, extrinsics, image_width, image_height):
        self.fx, self.fy, self.cx, self.cy, self.height = fx, fy, cx, cy, height
        self.E = np.asarray(extrinsics, np.float64).reshape(3, 3)
        self.W, self.H = image_width, image_height

    def convert_center_to_principal(self, p):           # camera.cc:148-158
        out = np.zeros(3)
        out[2] = p[2]
        o_bias = np.array([self.W * 0.5 - self.cx, self.H * 0.5 - self.cy])
        out[:2] = p[:2] + (np.eye(2) - rotation_matrix_2d(p[2])) @ o_bias
        return out

    def image_plane_to_camera(self, p):                 # camera.cc:160-175
        return np.array([p[0] / self.fx, p[1] / self.fy, p[2]])

    def camera_to_robot(self, p):                       # camera.cc:196-209
        return self.E @ np.array([self.height * p[0], self.height * p[1], p[2]])

    def image_plane_to_robot(self, p):                  # camera.cc:224-231
        return self.camera_to_robot(self.image_plane_to_camera(p))


class MapBuilderTracker:
    """compute_intermedium(image_f32) -> (fft_result, fft_polar); compute_pose(last_F, image_f32, last_P, P) -> (response[3], pose[3])
    are the CorrelationFlow calls (tracking mode), e.g. the C oracle's."""

    def __init__(self, camera, max_distance, max_angle, lower_response_thr, upper_response_thr, compute_intermedium, compute_pose):
        self.cam = camera
        self.kfs = (max_distance, max_angle, lower_response_thr, upper_response_thr)
        self.ci, self.cp = compute_intermedium, compute_pose
        self.init = False
        self.frame_id = 0
        self.keyframe = -1
        self.current_cf_pose = np.zeros(3)
        self.current_pose = np.zeros(3)

    def add_new_input(self, image_f32):
        """-> dict(tracked, inserted, keyframe, response, relative_pose, cf_pose, pose, distance); `inserted` is AddNewInput's return."""
        max_d, max_a, lo, hi = self.kfs
        F, P = self.ci(image_f32)                        # ComputeFFTResult :72-75
        fid = self.frame_id
        self.frame_id += 1
        if not self.init:                                # Initialize :86-97
            self.current_cf_pose = np.zeros(3)
            self.current_cf_real_pose = self.cam.image_plane_to_camera(self.current_cf_pose)
            self.current_pose = self.cam.camera_to_robot(self.current_cf_real_pose)
            self.distance = 0.0
            self.init = True
            self._update_intermedium(F, P, fid)
            return self._out(True, True, -1, np.zeros(3), np.zeros(3))
        kf = self.keyframe
        response, rel = self.cp(self.last_F, image_f32, self.last_P, P)          # Tracking :127-138
        response, rel = np.asarray(response, np.float64), np.asarray(rel, np.float64)
        rel = self.cam.convert_center_to_principal(rel)
        good = response[0] > lo and response[2] > lo
        inserted = False
        if good:
            self.current_cf_pose = compute_absolute_pose(self.last_cf_pose, rel)
            self.current_cf_real_pose = self.cam.image_plane_to_camera(self.current_cf_pose)
            # UpdateCurrentPose :118-125
            r0 = self.cam.image_plane_to_robot(self.last_cf_pose)
            r1 = self.cam.image_plane_to_robot(self.current_cf_pose)
            self.current_pose = compute_absolute_pose(self.last_pose, compute_relative_pose(r0, r1))
            # ComputeRelativeDA :157-166
            rc = self.cam.image_plane_to_camera(self.current_cf_pose - self.last_cf_pose)
            d, a = math.sqrt(rc[0] * rc[0] + rc[1] * rc[1]), abs(rc[2])
            c1, c2 = d > max_d, a > max_a
            c3 = lo < response[0] < hi
            c4 = lo < response[2] < hi
            inserted = c1 or c2 or c3 or c4              # :47-52
            if inserted:
                self.distance += d                       # :53
                self._update_intermedium(F, P, fid)      # :68
        return self._out(good, inserted, kf, response, rel)

    def _update_intermedium(self, F, P, fid):            # :99-106
        self.last_F, self.last_P = F, P
        self.last_cf_pose = self.current_cf_pose.copy()
        self.last_pose = self.current_pose.copy()
        self.keyframe = fid

    def _out(self, tracked, inserted, kf, response, rel):
        return dict(tracked=bool(tracked), inserted=bool(inserted), keyframe=kf, response=np.array(response, np.float64),
                    relative_pose=np.array(rel, np.float64), cf_pose=self.current_cf_pose.copy(), pose=self.current_pose.copy(),
                    distance=self.distance)
