// Minimal stand-in for boost::thread::physical_concurrency() (used at
// src/ndzip/cpu_factory.cc:9 of the reference when num_threads == 0).
#pragma once
#include <fstream>
#include <set>
#include <string>
#include <thread>
#include <utility>

namespace boost {

class thread {
  public:
    static unsigned physical_concurrency() {
        // distinct (physical id, core id) pairs from /proc/cpuinfo; falls back to logical CPUs
        std::ifstream cpuinfo("/proc/cpuinfo");
        std::set<std::pair<int, int>> cores;
        int phys = -1;
        std::string line;
        while (std::getline(cpuinfo, line)) {
            if (line.rfind("physical id", 0) == 0) {
                phys = std::stoi(line.substr(line.find(':') + 1));
            } else if (line.rfind("core id", 0) == 0) {
                cores.emplace(phys, std::stoi(line.substr(line.find(':') + 1)));
            }
        }
        if (!cores.empty()) return static_cast<unsigned>(cores.size());
        unsigned n = std::thread::hardware_concurrency();
        return n ? n : 1;
    }
};

}  // namespace boost
