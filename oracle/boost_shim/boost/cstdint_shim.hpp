// oracle/boost_shim -- a from-scratch, minimal stand-in for the handful of Boost headers that
// /root/reference/src/source.cc includes (source.cc:35-45).  TEST INFRASTRUCTURE ONLY: it lets
// the UNMODIFIED reference source compile in an image that has no Boost, so that its literal
// find_freq / find_pheno / apply_* / main() can be executed as the parity anchor and CPU
// baseline (oracle/Makefile -> oracle/_ref/).  Nothing here is Boost code; behaviour follows the
// published Boost documentation and is NOT pinned against a real Boost build.
#pragma once
#include <cstdint>
namespace boost { typedef std::int64_t int64_t; typedef std::uint32_t uint32_t; }
