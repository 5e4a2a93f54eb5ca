// CPU probe of the headless player's I/O (player/*.hpp), driven by tests/test_player_io.py.  Test infrastructure only.
#include "../../player/dataset_reader.hpp"
#include "../../player/trajectory_writer.hpp"
#include <cstdio>
#include <cstdlib>
#include <string>
using namespace xplayer;

static bool write_raw(const std::string &path, const GrayImage &img) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fwrite(img.data.data(), 1, img.data.size(), f);
    std::fclose(f);
    return true;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const std::string cmd = argv[1];
    if (cmd == "png" && argc == 4) {
        GrayImage img; std::string err;
        if (!read_gray_image(argv[2], img, err)) { std::printf("error %s\n", err.c_str()); return 1; }
        write_raw(argv[3], img);
        std::printf("%d %d\n", img.width, img.height);
        return 0;
    }
    if (cmd == "undistort" && argc == 16) {
        const std::string model = argv[2];
        const int w = std::atoi(argv[4]), h = std::atoi(argv[5]);
        double K[4], D[4];
        for (int i = 0; i < 4; ++i) { K[i] = std::atof(argv[6 + i]); D[i] = std::atof(argv[10 + i]); }
        std::vector<uint8_t> buf;
        if (!read_file(argv[14], buf) || buf.size() != (size_t)w * h) return 1;
        GrayImage src, dst;
        src.width = w; src.height = h; src.data = buf;
        Undistorter u;
        u.build(w, h, K, D, model);
        u.apply(src, dst);
        write_raw(argv[15], dst);
        return 0;
    }
    if (cmd == "events" && argc == 4) {
        auto reader = DatasetReader::create_reader(argv[2], argv[3]);
        if (!reader) { std::printf("no reader\n"); return 1; }
        DatasetReader::NextDataType t;
        while ((t = reader->next()) != DatasetReader::END) {
            if (t == DatasetReader::GYROSCOPE) { auto it = reader->read_gyroscope(); std::printf("G %.17g %.17g %.17g %.17g %.17g\n", it.first, it.second.data[0], it.second.data[1], it.second.data[2], it.second.timestamp); }
            else if (t == DatasetReader::ACCELEROMETER) { auto it = reader->read_accelerometer(); std::printf("A %.17g %.17g %.17g %.17g %.17g\n", it.first, it.second.data[0], it.second.data[1], it.second.data[2], it.second.timestamp); }
            else if (t == DatasetReader::CAMERA) {
                std::string err;
                auto it = reader->read_image(err);
                unsigned long long sum = 0;
                for (uint8_t v : it.second.data) sum += v;
                std::printf("C %.17g %d %d %llu %s\n", it.first, it.second.width, it.second.height, sum, err.c_str());
            }
        }
        return 0;
    }
    if (cmd == "write" && argc == 4) {
        XRSLAMPose p = {{0.1234567891234, -0.5, 0.25, 0.8}, {1.5, -2.25, 3.000000001}, 1403636580.838555574};
        if (std::string(argv[2]) == "tum") { TumTrajectoryWriter w(argv[3]); w.write_pose(p.timestamp, p); p.timestamp += 0.05; w.write_pose(p.timestamp, p); }
        else { CsvTrajectoryWriter w(argv[3]); w.write_pose(p.timestamp, p); }
        return 0;
    }
    return 2;
}
