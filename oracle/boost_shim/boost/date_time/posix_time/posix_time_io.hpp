#pragma once
#include "posix_time.hpp"
