"""CPU oracle for the xrslam VIO hot path -- TEST INFRASTRUCTURE ONLY.

This package is a float64 numpy (+ cv2 for the OpenCV image ops the reference
calls) restatement of the reference's per-frame front end and sliding-window
back end.  It exists so that tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs can check and time the CUDA path against
the reference's arithmetic.  Nothing in xrslam_b200/ (the product) may import
or call it.

Parity pinning status (see DESIGN.md "Oracle"):
  * front end (K1-K5): pinned against the reference's own known-answer test
    xrslam-test/test/src/test_feature_track.cpp:41,55,64 (164 / 161 counts on
    the two EuRoC V1_01 frames, reproduced to +-2 with cv2 4.13, the only
    OpenCV on this box) -- tests/test_oracle_frontend.py.
  * factors (K7, K8, K8b, K12 residual): pinned with the reference's
    finite-difference Jacobian recipe
    (xrslam/src/xrslam/estimation/ceres/cost_function_validator.h:22-26).
  * solver loop (K9-K11): Ceres 1.14 is NOT under /root/reference and is not
    installed; the DOGLEG trust-region loop is restated from the published
    1.14 algorithm.  PARITY UNPINNED for the loop itself (no reference test
    pins an iterate or a cost).
"""
