// NameIndex (plink_text.h): sharded, multi-threaded build above 4M names; a duplicate is reported as the
// reference's sequential std::map build would (source.cc:337-341): the name whose second occurrence comes first.
#include "../../simu_b200/host/plink_text.h"
#include <cstdio>
using namespace simu;
int main(){
  // duplicates: sequential semantics = report the name whose second occurrence comes first
  std::vector<std::string> store;
  for (long i=0;i<5000000;i++) store.push_back("rs"+std::to_string(i));
  store[4000000] = "rs77";       // second occurrence at 4000000
  store[3000000] = "rs123456";   // second occurrence at 3000000  -> reported
  std::vector<std::string_view> names(store.begin(), store.end());
  for (const char* t : {"1","8"}) {
    setenv("SIMU_B200_IO_THREADS", t, 1);
    NameIndex ix;
    try { ix.build(names); printf("threads %s: no error?!\n", t); return 1; }
    catch (std::exception& e) { printf("threads %s: %s\n", t, e.what()); if (std::string(e.what()) != "ERROR: duplicated variant name: rs123456") return 1; }
  }
  store[4000000] = "rs4000000"; store[3000000] = "rs3000000";
  std::vector<std::string_view> names2(store.begin(), store.end());
  setenv("SIMU_B200_IO_THREADS", "8", 1);
  NameIndex ix; ix.build(names2);
  if (ix.find("rs4999999") != 4999999 || ix.find("rs0") != 0 || ix.find("nope") != -1) return 1;
  printf("ok\n"); return 0;
}
