// crc32_fold_check -- csrc/host/crc32_fold.cc against zlib's crc32() for every length below 3000 at four alignments and for member-sized buffers.
#include <zlib.h>
#include <chrono>
#include <cstdio>
#include <vector>
#include "crc32_fold.hpp"
int main(){ std::vector<unsigned char> d((64<<20)+100); unsigned s=12345; for(auto&x:d){s=s*1664525u+1013904223u;x=(unsigned char)(s>>24);}
 long bad=0; for(size_t n=0;n<3000;++n) for(size_t off: {size_t(0),size_t(1),size_t(7),size_t(13)}) if(msnv::crc32_member(d.data()+off,n)!=(uint32_t)crc32(crc32(0L,nullptr,0),d.data()+off,(uInt)n)) {++bad; if(bad<5)printf("bad n=%zu off=%zu\n",n,off);}
 for(size_t n: {size_t(65280),size_t(65536),size_t(65535),size_t(1<<20)}) if(msnv::crc32_member(d.data()+3,n)!=(uint32_t)crc32(crc32(0L,nullptr,0),d.data()+3,(uInt)n)) ++bad;
 printf("mismatches: %ld\n",bad);
 for(int r=0;r<3;++r){auto t0=std::chrono::steady_clock::now(); unsigned c=0; for(size_t o=0;o+65280<=d.size();o+=65280) c^=msnv::crc32_member(d.data()+o,65280); auto t1=std::chrono::steady_clock::now(); printf("fold %.0f MB/s (%x)\n", (64<<20)/1e6/std::chrono::duration<double>(t1-t0).count(), c);} return bad!=0; }
