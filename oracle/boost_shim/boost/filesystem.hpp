// shim of <boost/filesystem.hpp>: exists()
#pragma once
#include <string>
#include <sys/stat.h>
namespace boost { namespace filesystem {
inline bool exists(const std::string &p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
} }
