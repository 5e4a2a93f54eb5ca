// Dataset readers of the headless player: EuRoC (euroc://<path to mav0>) and TUM-VI (tum://<path to mav0>)
// trees -- cam0/data.csv + cam0/data/*.png, imu0/data.csv -- turned into one time-ordered stream of
// gyroscope / accelerometer / camera events.  Behaviour follows xrslam-pc/player/src/IO/
// {dataset_reader.cpp:18-36, euroc_dataset_reader.cpp:3-124, euroc_dataset_reader.h:32-118,
// tum_dataset_reader.cpp:3-110, tum_dataset_reader.h:38-102}: timestamps are integer nanoseconds scaled
// by 1e-9, the camera time offset of the device yaml is added to image times, events are stable-sorted by
// time with the insertion order camera < gyroscope < accelerometer, images are undistorted in the reader.
// Unlike the reference's fscanf patterns, line ends may be CRLF or LF.
#pragma once
#include "../include/xrslam_compat.h"
#include "png_reader.hpp"
#include "undistort.hpp"
#include <algorithm>
#include <deque>
#include <fstream>
#include <memory>
#include <sstream>

namespace xplayer {

// ---- the few device-yaml keys the readers need (yaml_config.cpp:161-172) ----
struct CameraConfig {
    double K[4] = {0, 0, 0, 0};          // cam0.intrinsics fu fv cu cv
    double D[4] = {0, 0, 0, 0};          // cam0.distortion
    int distortion_flag = 0;             // cam0.camera_distortion_flag
    double time_offset = 0.0;            // cam0.time_offset
    bool ok = false;
};

inline bool yaml_numbers(const std::string &txt, const std::string &key, int n, double *out, size_t from) {
    size_t p = from;
    while (true) {      // whole-key match: "size" must not hit "subframe_size"
        p = txt.find(key + ":", p);
        if (p == std::string::npos) return false;
        if (p == 0 || !(isalnum((unsigned char)txt[p - 1]) || txt[p - 1] == '_')) break;
        p += key.size();
    }
    p += key.size() + 1;
    int got = 0;
    while (got < n && p < txt.size()) {
        while (p < txt.size() && !(isdigit((unsigned char)txt[p]) || txt[p] == '-' || txt[p] == '+' || txt[p] == '.')) {
            if (isalpha((unsigned char)txt[p])) return false;
            ++p;
        }
        if (p >= txt.size()) break;
        char *end = nullptr;
        out[got++] = std::strtod(txt.c_str() + p, &end);
        p = (size_t)(end - txt.c_str());
    }
    return got == n;
}

inline CameraConfig load_camera_config(const std::string &device_yaml) {
    CameraConfig c;
    std::ifstream f(device_yaml);
    if (!f) return c;
    std::stringstream ss;
    std::string line;
    while (std::getline(f, line)) {
        const size_t h = line.find('#');
        if (h != std::string::npos) line = line.substr(0, h);
        ss << line << '\n';
    }
    const std::string txt = ss.str();
    const size_t c0 = txt.find("cam0:");
    if (c0 == std::string::npos) return c;
    double v[4];
    if (!yaml_numbers(txt, "intrinsics", 4, c.K, c0)) return c;
    if (yaml_numbers(txt, "camera_distortion_flag", 1, v, c0)) c.distortion_flag = (int)v[0];
    yaml_numbers(txt, "distortion", 4, c.D, c0);          // whole-key match: not "camera_distortion_flag" / "distortion_model"
    if (yaml_numbers(txt, "time_offset", 1, v, c0)) c.time_offset = v[0];
    c.ok = true;
    return c;
}

// ---- csv files ----
struct CameraRow { double t; std::string filename; };
struct ImuRow { double t, w[3], a[3]; };

inline std::vector<std::string> read_lines(const std::string &path, bool &ok) {
    std::vector<std::string> out;
    std::ifstream f(path, std::ios::binary);
    ok = (bool)f;
    std::string line;
    while (ok && std::getline(f, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
        out.push_back(line);
    }
    return out;
}

inline bool load_camera_csv(const std::string &path, std::vector<CameraRow> &rows) {      // euroc_dataset_reader.h:39-59
    bool ok;
    const auto lines = read_lines(path, ok);
    rows.clear();
    for (size_t i = 1; i < lines.size(); ++i) {           // line 0 is the header
        const std::string &l = lines[i];
        const size_t c = l.find(',');
        if (c == std::string::npos || c == 0 || c + 1 >= l.size()) break;
        char *end = nullptr;
        double t = std::strtod(l.c_str(), &end);
        if (end != l.c_str() + c) break;
        t *= 1e-9;
        rows.push_back({t, l.substr(c + 1)});
    }
    return ok;
}

inline bool load_imu_csv(const std::string &path, std::vector<ImuRow> &rows) {            // euroc_dataset_reader.h:89-104
    bool ok;
    const auto lines = read_lines(path, ok);
    rows.clear();
    for (size_t i = 1; i < lines.size(); ++i) {
        ImuRow r;
        if (std::sscanf(lines[i].c_str(), "%lf,%lf,%lf,%lf,%lf,%lf,%lf", &r.t, &r.w[0], &r.w[1], &r.w[2], &r.a[0], &r.a[1], &r.a[2]) != 7) break;
        r.t *= 1e-9;
        rows.push_back(r);
    }
    return ok;
}

// ---- readers ----
class DatasetReader {
  public:
    enum NextDataType { AGAIN, GYROSCOPE, ACCELEROMETER, CAMERA, END };     // dataset_reader.h
    virtual ~DatasetReader() = default;
    NextDataType next() const { return all_data.empty() ? END : all_data.front().second; }
    std::pair<double, XRSLAMGyroscope> read_gyroscope() {
        if (gyroscope_data.empty()) return {};
        auto item = gyroscope_data.front();
        all_data.pop_front(); gyroscope_data.pop_front();
        return item;
    }
    std::pair<double, XRSLAMAcceleration> read_accelerometer() {
        if (accelerometer_data.empty()) return {};
        auto item = accelerometer_data.front();
        all_data.pop_front(); accelerometer_data.pop_front();
        return item;
    }
    // the undistorted 8-bit gray image of the next camera event; empty image + error text when the file cannot be decoded
    std::pair<double, GrayImage> read_image(std::string &err) {
        if (image_data.empty()) return {};
        auto item = image_data.front();
        all_data.pop_front(); image_data.pop_front();
        GrayImage raw, img;
        if (!read_gray_image(item.second, raw, err)) return {item.first, GrayImage()};
        if (config.distortion_flag && !skip_undistortion) {
            if (undistorter.width != raw.width || undistorter.height != raw.height)
                undistorter.build(raw.width, raw.height, config.K, config.D, model);
            undistorter.apply(raw, img);
        } else img = std::move(raw);
        image_width = img.width; image_height = img.height;
        return {item.first, std::move(img)};
    }
    void get_image_resolution(int &w, int &h) const { w = image_width; h = image_height; }
    bool skip_undistortion = false;      // the library undistorts on the GPU (player --device-undistort)
    size_t remaining() const { return all_data.size(); }

    static std::unique_ptr<DatasetReader> create_reader(const std::string &uri, const std::string &device_yaml);     // dataset_reader.cpp:18-36

  protected:
    DatasetReader(const std::string &root, const CameraConfig &cfg, const std::string &distortion_model, bool float_camera_params)
        : config(cfg), model(distortion_model), float_params(float_camera_params) {
        std::vector<CameraRow> cams;
        std::vector<ImuRow> imus;
        load_camera_csv(root + "/cam0/data.csv", cams);
        load_imu_csv(root + "/imu0/data.csv", imus);
        for (auto &c : cams) {
            image_data.emplace_back(c.t + cfg.time_offset, root + "/cam0/data/" + c.filename);
            all_data.emplace_back(c.t + cfg.time_offset, CAMERA);
        }
        for (auto &m : imus) {
            XRSLAMGyroscope g = {{m.w[0], m.w[1], m.w[2]}, m.t};
            gyroscope_data.emplace_back(m.t, g);
            all_data.emplace_back(m.t, GYROSCOPE);
            XRSLAMAcceleration a = {{m.a[0], m.a[1], m.a[2]}, m.t};
            accelerometer_data.emplace_back(m.t, a);
            all_data.emplace_back(m.t, ACCELEROMETER);
        }
        auto by_time = [](const auto &x, const auto &y) { return x.first < y.first; };
        std::stable_sort(all_data.begin(), all_data.end(), by_time);
        std::stable_sort(image_data.begin(), image_data.end(), by_time);
        std::stable_sort(gyroscope_data.begin(), gyroscope_data.end(), by_time);
        std::stable_sort(accelerometer_data.begin(), accelerometer_data.end(), by_time);
    }
    CameraConfig config;
    std::string model;
    bool float_params;
    Undistorter undistorter;
    std::deque<std::pair<double, NextDataType>> all_data;
    std::deque<std::pair<double, XRSLAMGyroscope>> gyroscope_data;
    std::deque<std::pair<double, XRSLAMAcceleration>> accelerometer_data;
    std::deque<std::pair<double, std::string>> image_data;
    int image_width = 0, image_height = 0;
};

// EuRoC: radtan through cv::undistort with float32 K / D (euroc_dataset_reader.cpp:62-72)
struct EurocDatasetReader : DatasetReader {
    EurocDatasetReader(const std::string &root, const CameraConfig &cfg) : DatasetReader(root, cfg, "radtan", true) {}
};
// TUM-VI: equidistant through ImageUndistorter with double K / D (tum_dataset_reader.cpp:62-76)
struct TumDatasetReader : DatasetReader {
    TumDatasetReader(const std::string &root, const CameraConfig &cfg) : DatasetReader(root, cfg, "equidistant", false) {}
};

inline std::unique_ptr<DatasetReader> DatasetReader::create_reader(const std::string &uri, const std::string &device_yaml) {
    const CameraConfig cfg = load_camera_config(device_yaml);
    if (!cfg.ok) return nullptr;
    auto scheme = [&](const char *s, std::string &path) {
        const size_t n = std::strlen(s);
        if (uri.size() >= n && uri.compare(0, n, s) == 0) { path = uri.substr(n); return true; }
        return false;
    };
    std::string path;
    if (scheme("euroc://", path)) return std::unique_ptr<DatasetReader>(new EurocDatasetReader(path, cfg));
    if (scheme("tum://", path)) return std::unique_ptr<DatasetReader>(new TumDatasetReader(path, cfg));
    return nullptr;
}

}  // namespace xplayer
