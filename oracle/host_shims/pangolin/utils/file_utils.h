/* oracle/host_shims -- TEST INFRASTRUCTURE ONLY.  Pangolin is not installed; GUI/src/Tools/RawLogReader.cpp uses exactly one function of it. */
#pragma once
#include <string>
#include <sys/stat.h>
namespace pangolin {
inline bool FileExists(const std::string& f) { struct stat st; return ::stat(f.c_str(), &st) == 0; }
}
