// Evidence for we_powf2_fast() (run by hand, ~10 s on 8 cores; result recorded in DESIGN.md):
//   g++ -O2 -std=c++17 -fno-builtin -ffp-contract=off -pthread tests/exhaustive_powf2.cpp -o /tmp/ex && /tmp/ex
//   floats=1006632961 mismatches=727320 max |low29-2^28| among mismatches=888704 (2^19.76)
// exhaustive: for every positive normal float x with x*x normal, compare full glibc-powf emulation with rn(x*x)
// and record how close (in units of the 29 discarded mantissa bits) the exact product is to a rounding midpoint
#include "../waterentropy_b200/csrc/we_numerics.cuh"
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <atomic>
int main(int argc,char**argv){
  int nt=8; std::vector<std::thread> th; std::vector<uint64_t> maxd(nt,0), nmis(nt,0), ntot(nt,0);
  uint32_t lo=we_f2u(0x1p-60f), hi=we_f2u(0x1p60f);
  for(int t=0;t<nt;++t) th.emplace_back([&,t]{
    WePowTables T{&we_h_log2tab[0][0], we_h_exp2tab};
    for(uint64_t u=(uint64_t)lo+t; u<=hi; u+=nt){ float x=we_u2f((uint32_t)u); double p=(double)x*(double)x; float fast=(float)p; float full=we_powf2(x,T); ntot[t]++;
      if(we_f2u(fast)!=we_f2u(full)){ nmis[t]++; uint64_t low=we_d2u(p)&((1ull<<29)-1); int64_t d=(int64_t)low-(1ll<<28); if(d<0)d=-d; if((uint64_t)d>maxd[t])maxd[t]=d; } }
  });
  for(auto&x:th)x.join(); uint64_t M=0,N=0,Tt=0; for(int t=0;t<nt;++t){ if(maxd[t]>M)M=maxd[t]; N+=nmis[t]; Tt+=ntot[t]; }
  printf("floats=%llu mismatches=%llu max |low29-2^28| among mismatches=%llu (2^%.2f)\n",(unsigned long long)Tt,(unsigned long long)N,(unsigned long long)M, log2((double)M));
}
