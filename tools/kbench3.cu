// tools/kbench3.cu -- development micro-benchmark: TWO time levels in ONE persistent "dataflow" launch of the
// single-level kernel body, scheduled so that the second level reads what the first one wrote while it is
// still in the 126 MB L2 (temporal blocking in L2, no new arithmetic kernel).
//
// Work item = (level l in {0,1}, row chunk i, z tile j).  CTAs draw tickets from a global counter; the ticket
// order interleaves level 0 chunk s with level 1 chunk s-LAG.  Level 1 chunk i waits (spin, acquire) until
// level 0 chunks i-1, i, i+1 are complete (RAW on the level-0 output rows +-4; the same set covers the WAR
// on the in-place overwrite of the older level).  Cross-check: bitwise vs two plain launches.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -o kbench3 kbench3.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct Args {
  const float* p; float* pp; const float* vdt;
  long long pitch; int ncol4, row0, row1, rows_per_cta;
  int lap_i0, lap_i1, lap_j0, lap_j1;
  int src_on, src_gi, src_j; float src_amp;
  float cz[9], cx[9];
};
__device__ __forceinline__ float4 ld4(const float* p){ return *reinterpret_cast<const float4*>(p);}
__device__ __forceinline__ float4 ldnc(const float* p){ return __ldg(reinterpret_cast<const float4*>(p));}
__device__ __forceinline__ float getk(const float4& v,int k){ return k==0?v.x:k==1?v.y:k==2?v.z:v.w; }
__device__ __forceinline__ float leap(float p,float pp,float t){ double d=__fma_rn(2.0,(double)p,-(double)pp); return __double2float_rn(__dadd_rn(d,(double)t)); }

// the production kernel's plain path, as a device function of (z tile bx, row chunk by)
template<bool VCG>
__device__ __forceinline__ void body(const Args& a,const float* P,float* PP,float amp,int bx,int by){
  constexpr int ORDER=8,H=4,W=9;
  int q=bx*blockDim.x+threadIdx.x;
  const bool act=q<a.ncol4;                                  // no early return: the caller has barriers after body()
  if(!act) q=a.ncol4-1;
  const int j0=q*4;
  const int rb=a.row0+by*a.rows_per_cta; const int re=min(rb+a.rows_per_cta,a.row1);
  const long long pitch=a.pitch;
  const bool ring = j0<a.lap_j0 || j0+4>a.lap_j1 || rb<a.lap_i0 || re>a.lap_i1;
  const bool near_src = a.src_on && a.src_j>=j0 && a.src_j<j0+4;
  const float* __restrict__ pc=P+j0+(long long)(rb-H)*pitch;
  float* __restrict__ ppc=PP+j0+(long long)rb*pitch;
  const float* __restrict__ vc=a.vdt+j0+(long long)rb*pitch;
  float4 w[W];
  #pragma unroll
  for(int s=0;s<2*H;s++){ w[s]=ld4(pc); pc+=pitch; }
  for(int left=re-rb; left>0; left-=W){
    #pragma unroll
    for(int u=0;u<W;u++){
      if(u<left){
        w[(u+2*H)%W]=ld4(pc);
        const float* ctr=pc-(long long)H*pitch;
        const float4 l=ld4(ctr-4), r=ld4(ctr+4), o=ld4(ppc), v=ldnc(vc);
        const float4 c4=w[(u+H)%W];
        const float za[12]={l.x,l.y,l.z,l.w,c4.x,c4.y,c4.z,c4.w,r.x,r.y,r.z,r.w};
        float lap[4];
        #pragma unroll
        for(int k2=0;k2<4;k2++){
          float az=__fmul_rn(za[k2],a.cz[0]); float ax=__fmul_rn(getk(w[u%W],k2),a.cx[0]);
          #pragma unroll
          for(int io=1;io<=ORDER;io++){ az=__fadd_rn(az,__fmul_rn(za[k2+io],a.cz[io])); ax=__fadd_rn(ax,__fmul_rn(getk(w[(u+io)%W],k2),a.cx[io])); }
          lap[k2]=__fadd_rn(az,ax);
        }
        if(ring){ const int lr=re-left+u; const bool rin=lr>=a.lap_i0&&lr<a.lap_i1;
          #pragma unroll
          for(int k2=0;k2<4;k2++) if(!rin||j0+k2<a.lap_j0||j0+k2>=a.lap_j1) lap[k2]=0.f; }
        float res[4];
        #pragma unroll
        for(int k2=0;k2<4;k2++) res[k2]=leap(getk(c4,k2),getk(o,k2),__fmul_rn(getk(v,k2),lap[k2]));
        if(near_src && re-left+u==a.src_gi){
          #pragma unroll
          for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res[k2]=__fadd_rn(res[k2],amp);
        }
        if(act) *reinterpret_cast<float4*>(ppc)=make_float4(res[0],res[1],res[2],res[3]);
        pc+=pitch; ppc+=pitch; vc+=pitch;
      }
    }
  }
}

__global__ void __launch_bounds__(256,4) k1(const __grid_constant__ Args a){ body<false>(a,a.p,a.pp,a.src_amp,blockIdx.x,blockIdx.y); }
// register-budget variants for 64-thread CTAs: MINB CTAs/SM -> 65536/(64*MINB) registers per thread
template<int MINB> __global__ void __launch_bounds__(64,MINB) k1r(const __grid_constant__ Args a){ body<false>(a,a.p,a.pp,a.src_amp,blockIdx.x,blockIdx.y); }
typedef void (*k1fn)(const Args);

struct DfArgs { Args a; float* bufA; float* bufB; float amp1, amp2; int nchunks, ntz, lag; unsigned* ticket; unsigned* done; int* err; };

// One CTA per work item; the item is a function of the linear block index, and CTAs are dispatched in block-index
// order, so every dependency of a CTA is already resident or finished when it starts to wait.
__global__ void __launch_bounds__(256,4) kdf(const __grid_constant__ DfArgs d){
  const int nch=d.nchunks, ntz=d.ntz, lag=d.lag;
  const int tile=blockIdx.x, y=blockIdx.y;                    // dispatch order: x fastest, then y
  int level,chunk;
  if(y<lag){ level=0; chunk=y; }
  else if(y<2*nch-lag){ const int r=y-lag; level=r&1; chunk=level?(r>>1):lag+(r>>1); }
  else { level=1; chunk=y-nch; }
  if(level==1){
    if(threadIdx.x==0){
      const long long t0=clock64();
      for(int c=max(chunk-1,0); c<=min(chunk+1,nch-1); c++)
        while(*((volatile unsigned*)d.done+c)<(unsigned)ntz){ if(clock64()-t0>(1LL<<31)){ atomicExch(d.err,1); break; } }
      __threadfence();
    }
    __syncthreads();
  }
  // level 0: t (in A) -> t+1 over t-1 (in B);  level 1: t+1 (in B) -> t+2 over t (in A)
  body<false>(d.a, level?d.bufB:d.bufA, level?d.bufA:d.bufB, level?d.amp2:d.amp1, tile, chunk);
  if(level==0){
    __syncthreads();
    if(threadIdx.x==0){ __threadfence(); atomicAdd(d.done+chunk,1u); }
  }
}

int main(int argc,char**argv){
  int n = argc>1?atoi(argv[1]):16384;
  const int G=16;
  long long pitch=((long long)n+4+31)/32*32; size_t rows=n+2*G; size_t elems=rows*pitch;
  float *A0,*B0,*V,*A,*B,*RA,*RB;
  cudaMalloc(&A0,elems*4); cudaMalloc(&B0,elems*4); cudaMalloc(&V,elems*4); cudaMalloc(&A,elems*4); cudaMalloc(&B,elems*4);
  cudaMalloc(&RA,elems*4); cudaMalloc(&RB,elems*4);
  std::vector<float> h(elems,0.f),x(elems),y(elems);
  srand(1);
  auto fill=[&](float* dd,int kind){ std::fill(h.begin(),h.end(),0.f);
    for(int i=0;i<n;i++) for(int j=0;j<n;j++) h[(size_t)(i+G)*pitch+j]= kind==2 ? 0.0625f+0.01f*((i*131+j)%7) : (rand()/(float)RAND_MAX-0.5f);
    cudaMemcpy(dd,h.data(),elems*4,cudaMemcpyHostToDevice); };
  fill(A0,0); fill(B0,1); fill(V,2);
  Args a; memset(&a,0,sizeof a); a.pitch=pitch; a.ncol4=(n+3)/4; a.row0=0;a.row1=n;
  a.lap_i0=4;a.lap_i1=n-4;a.lap_j0=4;a.lap_j1=n-4; a.src_on=1;a.src_gi=n/2;a.src_j=41;
  const float amp1=0.5f, amp2=-0.25f;
  for(int i=0;i<9;i++){int m=i<=4?i:8-i; a.cz[i]=0.01f*(m+1)*(m%2?1:-1); a.cx[i]=0.02f*(9-m)*(m%2?-1:1);}
  const long long o0=(long long)G*pitch; a.vdt=V+o0;
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int nsm; cudaDeviceGetAttribute(&nsm,cudaDevAttrMultiProcessorCount,0);
  const int NPAIR=4;   // timed: NPAIR pairs of levels back to back (sustained, like bench.py)
  auto cmp=[&](const float* d1,const float* d2){ cudaMemcpy(x.data(),d1,elems*4,cudaMemcpyDeviceToHost); cudaMemcpy(y.data(),d2,elems*4,cudaMemcpyDeviceToHost); return !memcmp(x.data(),y.data(),elems*4); };
  // ---- reference: plain launches
  for(int rpc: {32,16,8}){
    dim3 grid((a.ncol4+255)/256,(n+rpc-1)/rpc), block(256); Args r=a; r.rows_per_cta=rpc;
    float best=1e9;
    for(int rep=0;rep<3;rep++){
      cudaMemcpy(RA,A0,elems*4,cudaMemcpyDeviceToDevice); cudaMemcpy(RB,B0,elems*4,cudaMemcpyDeviceToDevice);
      cudaEventRecord(e0);
      for(int k=0;k<NPAIR;k++){
        r.p=RA+o0; r.pp=RB+o0; r.src_amp=amp1; k1<<<grid,block>>>(r);
        r.p=RB+o0; r.pp=RA+o0; r.src_amp=amp2; k1<<<grid,block>>>(r);
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms,e0,e1); if(rep>0&&ms<best) best=ms;
    }
    printf("K1 plain launches rpc=%2d            %.3f ms / 2 levels  %.1f Gpts/s\n",rpc,best/NPAIR,2.0*NPAIR*n*(double)n/(best*1e-3)/1e9); fflush(stdout);
  }
  // ---- register budget x geometry for 64-thread CTAs
  if(argc>2 && !strcmp(argv[2],"regs")){
    struct { const char* name; k1fn f; } rv[]={{"k1 (256,4) 64 regs",k1},{"k1r<16> 64 regs",k1r<16>},{"k1r<18> 56 regs",k1r<18>},{"k1r<20> 48 regs",k1r<20>},{"k1r<14> 72 regs",k1r<14>},{"k1r<12> 80 regs",k1r<12>}};
    for(auto& v: rv) for(int rpc: {7,14}){
      cudaFuncAttributes fa; cudaFuncGetAttributes(&fa,v.f);
      dim3 grid((a.ncol4+63)/64,(n+rpc-1)/rpc), block(64); Args r=a; r.rows_per_cta=rpc;
      float best=1e9;
      for(int rep=0;rep<3;rep++){
        cudaMemcpy(RA,A0,elems*4,cudaMemcpyDeviceToDevice); cudaMemcpy(RB,B0,elems*4,cudaMemcpyDeviceToDevice);
        cudaEventRecord(e0);
        for(int k=0;k<NPAIR;k++){
          r.p=RA+o0; r.pp=RB+o0; r.src_amp=amp1; v.f<<<grid,block>>>(r);
          r.p=RB+o0; r.pp=RA+o0; r.src_amp=amp2; v.f<<<grid,block>>>(r);
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms,e0,e1); if(rep>0&&ms<best) best=ms;
      }
      printf("%-22s regs=%3d spill=%4zu rpc=%2d  %.3f ms / 2 levels  %.1f Gpts/s\n",v.name,fa.numRegs,(size_t)fa.localSizeBytes,rpc,best/NPAIR,2.0*NPAIR*n*(double)n/(best*1e-3)/1e9); fflush(stdout);
    }
    return 0;
  }
  // ---- dataflow
  unsigned *ticket,*done; int* err; cudaMalloc(&ticket,4); cudaMalloc(&err,4); cudaMemset(err,0,4);
  cudaMalloc(&done,sizeof(unsigned)*NPAIR*(n/4+1));
  int occ=0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,kdf,256,0);
  for(int rpc: {8,16,32}) for(int lagrows: {64,128,256,384,512,1024}){
    DfArgs d; d.a=a; d.a.rows_per_cta=rpc; d.amp1=amp1; d.amp2=amp2; d.nchunks=(n+rpc-1)/rpc; d.ntz=(a.ncol4+255)/256; d.lag=lagrows/rpc; if(d.lag<2) d.lag=2;
    if(d.lag>=d.nchunks) continue;
    d.ticket=ticket; d.err=err; d.bufA=A+o0; d.bufB=B+o0; d.done=nullptr;
    float best=1e9;
    for(int rep=0;rep<3;rep++){
      cudaMemcpy(A,A0,elems*4,cudaMemcpyDeviceToDevice); cudaMemcpy(B,B0,elems*4,cudaMemcpyDeviceToDevice);
      cudaMemset(done,0,sizeof(unsigned)*NPAIR*(n/4+1));
      cudaEventRecord(e0);
      for(int k=0;k<NPAIR;k++){
        d.done=done+(size_t)k*(n/4+1);
        kdf<<<dim3(d.ntz,2*d.nchunks),256>>>(d);
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms,e0,e1); if(rep>0&&ms<best) best=ms;
    }
    cudaError_t er=cudaGetLastError(); if(er!=cudaSuccess){printf("ERR %s\n",cudaGetErrorString(er));return 1;}
    int herr=0; cudaMemcpy(&herr,err,4,cudaMemcpyDeviceToHost);
    int s1=cmp(A,RA), s2=cmp(B,RB);
    printf("DATAFLOW rpc=%2d lag=%4d rows occ=%d  %.3f ms / 2 levels  %.1f Gpts/s  same=%d,%d err=%d\n",rpc,d.lag*rpc,occ,best/NPAIR,2.0*NPAIR*n*(double)n/(best*1e-3)/1e9,s1,s2,herr); fflush(stdout);
  }
  return 0;
}
