// Headless player: the loop of xrslam-pc/player/src/main.cpp:56-169 without the viewer, linked against
// libxrvio.so through the reference's outer C ABI (include/xrslam_compat.h).  Usage, as in the reference:
//     xrvio_player -sc <slam.yaml> -dc <device.yaml> [--tum out.txt] [--csv out.csv] euroc://<mav0> | tum://<mav0>
// Differences a user of the reference player sees: no window, no license argument needed (-lc is accepted and
// ignored), the dataset reader takes the device yaml path instead of the YamlConfig* returned by XRSLAMCreate.
#include "../include/xrslam_compat.h"
#include "dataset_reader.hpp"
#include "trajectory_writer.hpp"
#include <cstdlib>
#include <iostream>

using namespace xplayer;

int main(int argc, char **argv) {
    std::string slam_config, device_config, license, csv_output, tum_output, input;
    // --device-undistort: hand the RAW camera frames to the library, which undistorts them on the GPU with the readers' own
    // arithmetic (xrvio_params.distortion_model, undistort_kernel) instead of on the CPU in the reader
    bool device_undistort = false;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&](std::string &dst) { if (i + 1 < argc) dst = argv[++i]; };
        if (a == "-sc" || a == "--slamconfig") val(slam_config);
        else if (a == "-dc" || a == "--deviceconfig") val(device_config);
        else if (a == "-lc" || a == "--license") val(license);
        else if (a == "--csv") val(csv_output);
        else if (a == "--tum") val(tum_output);
        else if (a == "-p" || a == "--play") {}
        else if (a == "--device-undistort") device_undistort = true;
        else input = a;
    }
    if (slam_config.empty() || device_config.empty() || input.empty()) {
        std::fprintf(stderr, "usage: %s -sc slam.yaml -dc device.yaml [--tum file] [--csv file] euroc://path | tum://path\n", argv[0]);
        return EXIT_FAILURE;
    }
    if (device_undistort) setenv("XRVIO_DEVICE_UNDISTORT", "1", 1);      // read by XRSLAMCreate (xrslam_shim.cu)
    void *config = nullptr;
    const int create_succ = XRSLAMCreate(slam_config.c_str(), device_config.c_str(), license.c_str(), "XRSLAM PC", &config);
    std::cout << "create SLAM success: " << create_succ << std::endl;
    if (!create_succ) return EXIT_FAILURE;        // the reference goes on and crashes later; fail loudly instead

    std::vector<std::unique_ptr<TrajectoryWriter>> outputs;
    try {
        if (!csv_output.empty()) outputs.emplace_back(new CsvTrajectoryWriter(csv_output));
        if (!tum_output.empty()) outputs.emplace_back(new TumTrajectoryWriter(tum_output));
    } catch (const std::exception &e) { std::fprintf(stderr, "%s\n", e.what()); XRSLAMDestroy(); return EXIT_FAILURE; }

    std::unique_ptr<DatasetReader> reader = DatasetReader::create_reader(input, device_config);
    if (reader && device_undistort) reader->skip_undistortion = true;
    if (!reader) { std::fprintf(stderr, "Cannot open \"%s\"\n", input.c_str()); XRSLAMDestroy(); return EXIT_FAILURE; }

    bool has_gyroscope = false, has_accelerometer = false;
    long frames = 0, tracked = 0;
    DatasetReader::NextDataType next_type;
    while ((next_type = reader->next()) != DatasetReader::END) {
        switch (next_type) {
        case DatasetReader::AGAIN: continue;
        case DatasetReader::GYROSCOPE: {
            has_gyroscope = true;
            auto item = reader->read_gyroscope();
            XRSLAMPushSensorData(XRSLAM_SENSOR_GYROSCOPE, &item.second);
        } break;
        case DatasetReader::ACCELEROMETER: {
            has_accelerometer = true;
            auto item = reader->read_accelerometer();
            XRSLAMPushSensorData(XRSLAM_SENSOR_ACCELERATION, &item.second);
        } break;
        case DatasetReader::CAMERA: {
            std::string err;
            auto item = reader->read_image(err);
            if (item.second.empty()) { std::fprintf(stderr, "image at t = %.9f: %s\n", item.first, err.c_str()); XRSLAMDestroy(); return EXIT_FAILURE; }
            XRSLAMImage image;
            image.camera_id = 0;
            image.timeStamp = item.first;
            image.ext = nullptr;
            image.data = item.second.data.data();
            image.channel = 1;
            image.stride = item.second.width;
            XRSLAMPushSensorData(XRSLAM_SENSOR_CAMERA, &image);
            frames++;
            if (has_accelerometer && has_gyroscope) {
                XRSLAMRunOneFrame();
                XRSLAMState state;
                XRSLAMGetResult(XRSLAM_RESULT_STATE, &state);
                if (state == XRSLAM_STATE_TRACKING_SUCCESS) {
                    XRSLAMPose pose_b;
                    XRSLAMGetResult(XRSLAM_RESULT_BODY_POSE, &pose_b);
                    if (pose_b.timestamp > 0) {
                        tracked++;
                        for (auto &output : outputs) output->write_pose(pose_b.timestamp, pose_b);
                    }
                }
            }
        } break;
        default: break;
        }
    }
    outputs.clear();
    XRSLAMDestroy();
    std::cout << "frames: " << frames << " tracked poses: " << tracked << std::endl;
    return EXIT_SUCCESS;
}
