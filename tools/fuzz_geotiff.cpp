// Mutation fuzzer for librfsi_io's reader: corrupted files must fail cleanly, never crash.
//   g++ -O1 -g -fsanitize=address,undefined -std=c++17 -pthread -Iinclude tools/fuzz_geotiff.cpp \
//       rfsi_b200/csrc/geotiff.cpp -lz -o /tmp/fuzz_geotiff && /tmp/fuzz_geotiff 20000 1 seed*.tif
// (iterations per seed file, RNG seed, seed files; writes its candidates to /tmp/fuzz_geotiff_cur.tif).
// Round 1: 1.4 M mutants of 13 seed files (every layout of tests/test_geotiff.py + three of the
// reference's LZW files) after the 64-bit offset wrap-around it found was fixed: no ASan/UBSan report.
#include "rfsi_io.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <random>
static std::vector<unsigned char> slurp(const char* p){FILE*f=fopen(p,"rb");std::vector<unsigned char> v;if(!f)return v;fseek(f,0,SEEK_END);long n=ftell(f);fseek(f,0,SEEK_SET);v.resize(n);fread(v.data(),1,n,f);fclose(f);return v;}
int main(int argc,char**argv){
  int iters=atoi(argv[1]); std::mt19937_64 rng(atoi(argv[2]));
  long ok=0,bad=0;
  for(int a=3;a<argc;++a){
    auto base=slurp(argv[a]); if(base.empty()) continue;
    for(int it=0;it<iters;++it){
      auto v=base;
      int nm=1+rng()%8;
      for(int k=0;k<nm;++k){
        int mode=rng()%4; size_t pos=rng()%v.size();
        // bias towards the header / directory region
        if(rng()%2){ size_t ifd=0; memcpy(&ifd,&v[4],4); ifd&=0xffffffff; if(ifd<v.size()) pos=ifd+rng()%std::min<size_t>(v.size()-ifd,400); if(rng()%4==0) pos=rng()%16; }
        if(mode==0) v[pos]^=1u<<(rng()%8); else if(mode==1) v[pos]=rng(); else if(mode==2&&pos+4<=v.size()){unsigned x=rng()%3==0?0xffffffffu:(unsigned)rng(); memcpy(&v[pos],&x,4);} else if(mode==3 && rng()%8==0) v.resize(std::max<size_t>(1,pos));
      }
      FILE*f=fopen("/tmp/fuzz_geotiff_cur.tif","wb"); fwrite(v.data(),1,v.size(),f); fclose(f);
      rfsi_tiff_t* t=nullptr;
      if(rfsi_tiff_open("/tmp/fuzz_geotiff_cur.tif",&t)!=0){++bad;continue;}
      rfsi_tiff_info_t i; rfsi_tiff_get_info(t,&i);
      size_t cell = (i.dtype==0)?8:(i.dtype==1||i.dtype==2||i.dtype==6)?4:(i.dtype==3||i.dtype==4)?2:1;
      double bytes=(double)i.ncol*i.nrow*i.nlayers*cell;
      if(bytes>0 && bytes<3e8){ std::vector<unsigned char> out((size_t)bytes); if(rfsi_tiff_read(t,out.data(),out.size(),2)==0)++ok; else ++bad; } else ++bad;
      rfsi_tiff_close(t);
    }
  }
  printf("ok %ld rejected %ld\n",ok,bad); return 0; }
