/* xrslam_compat.h -- the subset of the reference's outer C ABI (xrslam-interface/include/XRSLAM.h:201-229)
 * that libxrvio.so exports so that xrslam-pc/player and xrslam-ros can link it in place of libxrslam:
 *     XRSLAMCreate / XRSLAMPushSensorData / XRSLAMRunOneFrame / XRSLAMGetResult / XRSLAMDestroy.
 * Struct names, field order and enum values restate that header (XRSLAM.h:16-199) so that callers compiled
 * against the reference's header are binary compatible.  One process-wide instance, as in the reference
 * (xrslam-interface/src/XRSLAMManager.cpp:6-9).  Extension: XRSLAMExtPushBootstrapState, see INTEGRATION.md. */
#ifndef XRSLAM_COMPAT_H
#define XRSLAM_COMPAT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum XRSLAMSensorType {
    XRSLAM_SENSOR_CAMERA = 0, XRSLAM_SENSOR_DEPTH_CAMERA, XRSLAM_SENSOR_ACCELERATION, XRSLAM_SENSOR_GYROSCOPE,
    XRSLAM_SENSOR_GRAVITY, XRSLAM_SENSOR_ROTATION_VECTOR, XRSLAM_SENSOR_UNKNOWN
} XRSLAMSensorType;

typedef struct XRSLAMImageExtension { double exposure_time, default_focus_distance, focal_length, focus_distance; } XRSLAMImageExtension;
typedef struct XRSLAMImage {
    unsigned char *data;
    double timeStamp;
    int stride, camera_id, channel;
    XRSLAMImageExtension *ext;
} XRSLAMImage;
typedef struct XRSLAMAcceleration { double data[3]; double timestamp; } XRSLAMAcceleration;
typedef struct XRSLAMGyroscope { double data[3]; double timestamp; } XRSLAMGyroscope;

typedef enum XRSLAMResultType {
    XRSLAM_RESULT_BODY_POSE = 0, XRSLAM_RESULT_CAMERA_POSE, XRSLAM_RESULT_STATE, XRSLAM_RESULT_LANDMARKS, XRSLAM_RESULT_FEATURES,
    XRSLAM_RESULT_BIAS, XRSLAM_RESULT_DEBUG_LOGS, XRSLAM_RESULT_VERSION, XRSLAM_RESULT_UNKNOWN, XRSLAM_INFO_INTRINSICS
} XRSLAMResultType;

typedef struct XRSLAMPose { double quaternion[4]; double translation[3]; double timestamp; } XRSLAMPose;
typedef struct XRSLAMIntrinsics { double fx, fy, cx, cy; } XRSLAMIntrinsics;
typedef enum XRSLAMState { XRSLAM_STATE_INITIALIZING, XRSLAM_STATE_TRACKING_SUCCESS, XRSLAM_STATE_TRACKING_FAIL } XRSLAMState;
typedef struct XRSLAMLandmark { double x, y, z; } XRSLAMLandmark;
typedef struct XRSLAMLandmarks { XRSLAMLandmark *landmarks; int num_landmarks; } XRSLAMLandmarks;
typedef struct XRSLAMBias { double data[3]; } XRSLAMBias;
typedef struct XRSLAMIMUBias { XRSLAMBias acc_bias, gyr_bias; } XRSLAMIMUBias;
typedef struct XRSLAMStringOutput { int str_length; char *data; } XRSLAMStringOutput;

int XRSLAMCreate(const char *slam_config_path, const char *device_config_path, const char *license_path, const char *product_name,
                 void **config);
void XRSLAMPushSensorData(XRSLAMSensorType sensor_type, void *sensor_data);
void XRSLAMRunOneFrame(void);
void XRSLAMGetResult(XRSLAMResultType result_type, void *result_data);
void XRSLAMDestroy(void);
/* Extension (not in the reference): ground-truth / externally estimated state of a frame time, consumed by the
 * bootstrap that stands in for the reference's Initializer. */
void XRSLAMExtPushBootstrapState(double t, const double q_xyzw[4], const double p[3], const double v[3]);

#ifdef __cplusplus
}
#endif
#endif
