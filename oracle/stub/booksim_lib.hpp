// Stand-in for the Booksim2 fork's library interface. Signatures taken from the
// reference call sites (src/chip.cpp:97-98, src/schedule.cpp:131-159,
// src/main.cpp:75). Calling the cycle-accurate model through this stub returns 0.
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>
#include "booksim_config.hpp"
inline BookSimConfig booksim_load_config(std::vector<std::string>) { return {}; }
inline void booksim_create_processing_event(int, std::pair<std::string, int>, std::pair<int, int>, double) {}
inline void booksim_create_spike_event(int, std::pair<std::string, int>, std::pair<int, int>, std::pair<int, int>, double, double) {}
inline double booksim_run(std::shared_ptr<BookSimConfig>) { return 0.0; }
inline void booksim_close() {}
