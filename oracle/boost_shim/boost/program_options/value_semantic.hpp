#pragma once
#include "../program_options.hpp"
