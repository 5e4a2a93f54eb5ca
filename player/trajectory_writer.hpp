// Trajectory writers of the player (xrslam-pc/player/src/IO/trajectory_writer.h:30-76): the CSV and TUM
// line formats are the reference's printf formats, so evo / the reference's tooling read the files unchanged.
#pragma once
#include "../include/xrslam_compat.h"
#include <cstdio>
#include <stdexcept>
#include <string>

namespace xplayer {

class TrajectoryWriter {
  public:
    virtual ~TrajectoryWriter() = default;
    virtual void write_pose(double t, const XRSLAMPose &pose) = 0;
};

class FileTrajectoryWriter : public TrajectoryWriter {
  protected:
    FILE *file;
    const char *fmt;
  public:
    FileTrajectoryWriter(const std::string &filename, const char *line_format) : file(std::fopen(filename.c_str(), "w")), fmt(line_format) {
        if (!file) throw std::runtime_error("Cannot open file " + filename);
    }
    ~FileTrajectoryWriter() override { std::fflush(file); std::fclose(file); }
    void write_pose(double t, const XRSLAMPose &pose) override {
        std::fprintf(file, fmt, t, pose.translation[0], pose.translation[1], pose.translation[2], pose.quaternion[0], pose.quaternion[1],
                     pose.quaternion[2], pose.quaternion[3]);
        std::fflush(file);
    }
};

struct CsvTrajectoryWriter : FileTrajectoryWriter {
    explicit CsvTrajectoryWriter(const std::string &f) : FileTrajectoryWriter(f, "%.18e,%.9e,%.9e,%.9e,%.7e,%.7e,%.7e,%.7e\n") {}
};
struct TumTrajectoryWriter : FileTrajectoryWriter {
    explicit TumTrajectoryWriter(const std::string &f) : FileTrajectoryWriter(f, "%.18e %.9e %.9e %.9e %.7e %.7e %.7e %.7e\n") {}
};

}  // namespace xplayer
