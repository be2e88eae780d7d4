// tools/kbench.cu -- development micro-benchmark: variants of the fused step kernel's plain path
// on a 16384^2 grid, timed with CUDA events, cross-checked bitwise against variant 0.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -o kbench kbench.cu
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
__device__ __forceinline__ float4 ldcs(const float* p){ float4 r; asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x),"=f"(r.y),"=f"(r.z),"=f"(r.w) : "l"(p)); return r;}
__device__ __forceinline__ void stcs(float* p, float4 v){ asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p),"f"(v.x),"f"(v.y),"f"(v.z),"f"(v.w) : "memory"); }
__device__ __forceinline__ float getk(const float4& v,int k){ return k==0?v.x:k==1?v.y:k==2?v.z:v.w; }
__device__ __forceinline__ float leap(float p,float pp,float t){ double d=__fma_rn(2.0,(double)p,-(double)pp); return __double2float_rn(__dadd_rn(d,(double)t)); }

struct Row { float4 wn,l,r,o,v; };

// PF: explicit prefetch of next row's operands; HINT: streaming cache hints on pp/vdt
template<int PF,int HINT,int MINB>
__global__ void __launch_bounds__(256,MINB) k(const __grid_constant__ Args a){
  constexpr int ORDER=8,H=4,W=9;
  const int q=blockIdx.x*blockDim.x+threadIdx.x; if(q>=a.ncol4) return;
  const int j0=q*4;
  const int rb=a.row0+blockIdx.y*a.rows_per_cta; const int re=min(rb+a.rows_per_cta,a.row1); if(rb>=re) return;
  const long long pitch=a.pitch;
  unsigned mlap=0;
  #pragma unroll
  for(int k2=0;k2<4;k2++) if(j0+k2>=a.lap_j0&&j0+k2<a.lap_j1) mlap|=1u<<k2;
  const float* __restrict__ pc=a.p+j0+(long long)(rb-H)*pitch;
  float* __restrict__ ppc=a.pp+j0+(long long)rb*pitch;
  const float* __restrict__ vc=a.vdt+j0+(long long)rb*pitch;
  float4 w[W];
  #pragma unroll
  for(int s=0;s<2*H;s++){ w[s]=ld4(pc); pc+=pitch; }
  auto loadrow=[&](const float* pcur,const float* ppcur,const float* vcur){ Row r; r.wn=ld4(pcur); const float* ctr=pcur-(long long)H*pitch; r.l=ld4(ctr-4); r.r=ld4(ctr+4);
      if(HINT){ r.o=ldcs(ppcur); r.v=ldcs(vcur);} else { r.o=ld4(ppcur); r.v=ldnc(vcur);} return r; };
  Row cur; if(PF) cur=loadrow(pc,ppc,vc);
  for(int r=rb;r<re;r+=W){
    #pragma unroll
    for(int u=0;u<W;u++){
      const int lr=r+u;
      if(lr<re){
        Row nx;
        if(PF){ if(lr+1<re) nx=loadrow(pc+pitch,ppc+pitch,vc+pitch); }
        else cur=loadrow(pc,ppc,vc);
        w[(u+2*H)%W]=cur.wn;
        const float4 c4=w[(u+H)%W];
        const float za[12]={cur.l.x,cur.l.y,cur.l.z,cur.l.w,c4.x,c4.y,c4.z,c4.w,cur.r.x,cur.r.y,cur.r.z,cur.r.w};
        const unsigned ml=(lr>=a.lap_i0&&lr<a.lap_i1)?mlap:0u;
        float res[4];
        #pragma unroll
        for(int k2=0;k2<4;k2++){
          float az=__fmul_rn(za[k2],a.cz[0]); float ax=__fmul_rn(getk(w[u%W],k2),a.cx[0]);
          #pragma unroll
          for(int io=1;io<=ORDER;io++){ az=__fadd_rn(az,__fmul_rn(za[k2+io],a.cz[io])); ax=__fadd_rn(ax,__fmul_rn(getk(w[(u+io)%W],k2),a.cx[io])); }
          float lap=__fadd_rn(az,ax); if(!((ml>>k2)&1u)) lap=0.f;
          res[k2]=leap(getk(c4,k2),getk(cur.o,k2),__fmul_rn(getk(cur.v,k2),lap));
        }
        if(a.src_on && lr==a.src_gi && a.src_j>=j0 && a.src_j<j0+4){
          #pragma unroll
          for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res[k2]=__fadd_rn(res[k2],a.src_amp);
        }
        float4 out=make_float4(res[0],res[1],res[2],res[3]);
        if(HINT) stcs(ppc,out); else *reinterpret_cast<float4*>(ppc)=out;
        pc+=pitch; ppc+=pitch; vc+=pitch;
        if(PF) cur=nx;
      }
    }
  }
}


// SHFL variant: z neighbours from the neighbouring lanes' centre-row registers; each warp computes 30 float4
// columns (lanes 1..30), lanes 0 and 31 only carry the halo columns.  Fresh loads per row: wn, o, v -- all
// consumed at the END of the row's arithmetic, so their latency overlaps the ~150 FP instructions before.
template<int MINB, int NCHK>
__global__ void __launch_bounds__(256,MINB) ks(const __grid_constant__ Args a){
  constexpr int ORDER=8,H=4,W=9;
  const int lane=threadIdx.x&31, warp=threadIdx.x>>5;
  const int wpb=blockDim.x>>5;
  const int q=(blockIdx.x*wpb+warp)*30+lane-1;        // float4 column, may be -1 or >= ncol4 on halo lanes
  if((blockIdx.x*wpb+warp)*30>=a.ncol4) return;         // whole warp out of range
  const bool comp = lane>=1 && lane<=30 && q<a.ncol4;
  const int j0=q*4;
  const int rb=a.row0+blockIdx.y*a.rows_per_cta; const int re=min(rb+a.rows_per_cta,a.row1); if(rb>=re) return;
  const long long pitch=a.pitch;
  const bool ring = j0<a.lap_j0 || j0+4>a.lap_j1 || rb<a.lap_i0 || re>a.lap_i1;
  const bool near_src = a.src_on && a.src_j>=j0 && a.src_j<j0+4;
  const float* __restrict__ pc=a.p+j0+(long long)(rb-H)*pitch;
  float* __restrict__ ppc=a.pp+j0+(long long)rb*pitch;
  const float* __restrict__ vc=a.vdt+j0+(long long)rb*pitch;
  float4 w[W];
  #pragma unroll
  for(int s=0;s<2*H;s++){ w[s]=ld4(pc); pc+=pitch; }
  for(int left=re-rb; left>0; left-=W){
    #pragma unroll
    for(int u=0;u<W;u++){
      if(u<left){
        float4 wn=ld4(pc);
        float4 o=make_float4(0,0,0,0), v=o;
        if(comp){ o=ld4(ppc); v=ldnc(vc); }
        const float4 c4=w[(u+H)%W];
        float4 l,r;
        l.x=__shfl_up_sync(0xffffffffu,c4.x,1); l.y=__shfl_up_sync(0xffffffffu,c4.y,1); l.z=__shfl_up_sync(0xffffffffu,c4.z,1); l.w=__shfl_up_sync(0xffffffffu,c4.w,1);
        r.x=__shfl_down_sync(0xffffffffu,c4.x,1); r.y=__shfl_down_sync(0xffffffffu,c4.y,1); r.z=__shfl_down_sync(0xffffffffu,c4.z,1); r.w=__shfl_down_sync(0xffffffffu,c4.w,1);
        const float za[12]={l.x,l.y,l.z,l.w,c4.x,c4.y,c4.z,c4.w,r.x,r.y,r.z,r.w};
        float lap[4];
        #pragma unroll
        for(int k2=0;k2<4;k2++){
          float az=__fmul_rn(za[k2],a.cz[0]); float ax=__fmul_rn(getk(w[u%W],k2),a.cx[0]);
          #pragma unroll
          for(int io=1;io<ORDER;io++){ az=__fadd_rn(az,__fmul_rn(za[k2+io],a.cz[io])); ax=__fadd_rn(ax,__fmul_rn(getk(w[(u+io)%W],k2),a.cx[io])); }
          az=__fadd_rn(az,__fmul_rn(za[k2+ORDER],a.cz[ORDER]));
          ax=__fadd_rn(ax,__fmul_rn(getk(wn,k2),a.cx[ORDER]));     // the freshly loaded row is the LAST tap
          lap[k2]=__fadd_rn(az,ax);
        }
        w[(u+2*H)%W]=wn;
        if(ring){ const int lr=re-left+u; const bool rin=lr>=a.lap_i0&&lr<a.lap_i1;
          #pragma unroll
          for(int k2=0;k2<4;k2++) if(!rin||j0+k2<a.lap_j0||j0+k2>=a.lap_j1) lap[k2]=0.f; }
        float res[4];
        #pragma unroll
        for(int k2=0;k2<4;k2++) res[k2]=leap(getk(c4,k2),getk(o,k2),__fmul_rn(getk(v,k2),lap[k2]));
        if(near_src && re-left+u==a.src_gi){
          #pragma unroll
          for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res[k2]=__fadd_rn(res[k2],a.src_amp);
        }
        if(comp) *reinterpret_cast<float4*>(ppc)=make_float4(res[0],res[1],res[2],res[3]);
        pc+=pitch; ppc+=pitch; vc+=pitch;
      }
    }
  }
}

__device__ __forceinline__ void pf_l2(const void* p){ asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
__device__ __forceinline__ void pf_l1(const void* p){ asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
// lean loop + cache prefetch D rows ahead (no registers): LVL 2 = L2, 1 = L1
template<int MINB,int D,int LVL>
__global__ void __launch_bounds__(256,MINB) kp(const __grid_constant__ Args a){
  constexpr int ORDER=8,H=4,W=9;
  const int q=blockIdx.x*blockDim.x+threadIdx.x; if(q>=a.ncol4) return;
  const int j0=q*4;
  const int rb=a.row0+blockIdx.y*a.rows_per_cta; const int re=min(rb+a.rows_per_cta,a.row1); if(rb>=re) return;
  const long long pitch=a.pitch;
  const bool ring = j0<a.lap_j0 || j0+4>a.lap_j1 || rb<a.lap_i0 || re>a.lap_i1;
  const bool near_src = a.src_on && a.src_j>=j0 && a.src_j<j0+4;
  const bool pfl = (threadIdx.x & 7)==0;   // one lane per 128-byte line
  const float* __restrict__ pc=a.p+j0+(long long)(rb-H)*pitch;
  float* __restrict__ ppc=a.pp+j0+(long long)rb*pitch;
  const float* __restrict__ vc=a.vdt+j0+(long long)rb*pitch;
  float4 w[W];
  #pragma unroll
  for(int s=0;s<2*H;s++){ w[s]=ld4(pc); pc+=pitch; }
  if(pfl){
    #pragma unroll
    for(int d=0;d<D;d++){ if(LVL==2){pf_l2(pc+d*pitch);pf_l2(ppc+d*pitch);pf_l2(vc+d*pitch);} else {pf_l1(pc+d*pitch);pf_l1(ppc+d*pitch);pf_l1(vc+d*pitch);} }
  }
  for(int left=re-rb; left>0; left-=W){
    #pragma unroll
    for(int u=0;u<W;u++){
      if(u<left){
        if(pfl){ if(LVL==2){pf_l2(pc+D*pitch);pf_l2(ppc+D*pitch);pf_l2(vc+D*pitch);} else {pf_l1(pc+D*pitch);pf_l1(ppc+D*pitch);pf_l1(vc+D*pitch);} }
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
          for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res[k2]=__fadd_rn(res[k2],a.src_amp);
        }
        *reinterpret_cast<float4*>(ppc)=make_float4(res[0],res[1],res[2],res[3]);
        pc+=pitch; ppc+=pitch; vc+=pitch;
      }
    }
  }
}

typedef void (*kfn)(const Args);
struct Var { const char* name; kfn f; int minb; int shfl; };
#define V(PF,HINT,MINB) {"PF" #PF "_H" #HINT "_B" #MINB, k<PF,HINT,MINB>, MINB, 0}
#define VS(MINB) {"SHFL_B" #MINB, ks<MINB,0>, MINB, 1}
#define VP(MINB,D,LVL) {"PFL" #LVL "_D" #D "_B" #MINB, kp<MINB,D,LVL>, MINB, 0}

int main(int argc,char**argv){
  int n = argc>1?atoi(argv[1]):16384;
  long long pitch=((long long)n+4+31)/32*32; size_t rows=n+10; size_t elems=rows*pitch;
  float *p,*pp0,*pp,*v,*ref;
  cudaMalloc(&p,elems*4); cudaMalloc(&pp0,elems*4); cudaMalloc(&pp,elems*4); cudaMalloc(&v,elems*4); cudaMalloc(&ref,elems*4);
  std::vector<float> h(elems); 
  srand(1); for(size_t i=0;i<elems;i++) h[i]=(rand()/(float)RAND_MAX-0.5f); cudaMemcpy(p,h.data(),elems*4,cudaMemcpyHostToDevice);
  for(size_t i=0;i<elems;i++) h[i]=(rand()/(float)RAND_MAX-0.5f); cudaMemcpy(pp0,h.data(),elems*4,cudaMemcpyHostToDevice);
  for(size_t i=0;i<elems;i++) h[i]=6.25f+ (i%7); cudaMemcpy(v,h.data(),elems*4,cudaMemcpyHostToDevice);
  Args a; memset(&a,0,sizeof a); a.p=p+5*pitch; a.vdt=v+5*pitch; a.pitch=pitch; a.ncol4=(n+3)/4; a.row0=0;a.row1=n;
  a.lap_i0=4;a.lap_i1=n-4;a.lap_j0=4;a.lap_j1=n-4; a.src_on=1;a.src_gi=n/2;a.src_j=40;a.src_amp=0.5f;
  for(int i=0;i<9;i++){a.cz[i]=0.01f*(i+1)*(i%2?1:-1); a.cx[i]=0.02f*(9-i)*(i%2?-1:1);}
  Var vars[]={VP(4,0,2),VP(4,4,2),VP(4,8,2),VP(4,16,2),VP(4,4,1),VP(4,8,1),VP(3,8,2)};
  int nsm; cudaDeviceGetAttribute(&nsm,cudaDevAttrMultiProcessorCount,0);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bool have_ref=false;
  for(auto& vr:vars){
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa,vr.f);
    for(int nt: {256}) for(int rpc: {32,64,128,148}){
      int occ=0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,vr.f,nt,0);
      int cols_per_blk = vr.shfl ? (nt/32)*30 : nt; dim3 grid((a.ncol4+cols_per_blk-1)/cols_per_blk,(n+rpc-1)/rpc), block(nt); a.rows_per_cta=rpc;
      float best=1e9;
      for(int rep=0;rep<4;rep++){
        cudaMemcpy(pp,pp0,elems*4,cudaMemcpyDeviceToDevice); a.pp=pp+5*pitch;
        cudaEventRecord(e0); vr.f<<<grid,block>>>(a); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms,e0,e1); if(rep>0&&ms<best) best=ms;
      }
      cudaError_t err=cudaGetLastError(); if(err!=cudaSuccess){printf("ERR %s\n",cudaGetErrorString(err));return 1;}
      if(!have_ref){ cudaMemcpy(ref,pp,elems*4,cudaMemcpyDeviceToDevice); have_ref=true; }
      // compare
      std::vector<float> x(elems),y(elems); 
      static int checked=0; int same=-1;
      if(checked<40){ cudaMemcpy(x.data(),pp,elems*4,cudaMemcpyDeviceToHost); cudaMemcpy(y.data(),ref,elems*4,cudaMemcpyDeviceToHost); same=!memcmp(x.data(),y.data(),elems*4); checked++; }
      double gbs=16.0*n*(double)n/(best*1e-3)/1e9;
      printf("%-12s regs=%3d nt=%3d rpc=%3d occ=%d  %.3f ms  %.0f GB/s  %.1f Gpts/s same=%d\n",vr.name,fa.numRegs,nt,rpc,occ,best,gbs,gbs/16,same);
      fflush(stdout);
    }
  }
  return 0;
}
