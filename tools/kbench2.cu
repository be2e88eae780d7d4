// tools/kbench2.cu -- development micro-benchmark: TWO time levels per pass (temporal blocking).
//
// One warp is an independent streaming unit: 32 float4 columns of level t+1 (lanes 0 and 31 are the
// z halo), 30 float4 columns of level t+2.  Streaming along x with a 4-row skew between the levels:
// at iteration rA the warp computes u(t+1)[rA] from the register window of u(t) (z neighbours via
// L1 loads), pushes it into the register window of u(t+1), and computes u(t+2)[rA-4] from that
// window (z neighbours through warp shuffles).  Inputs A=u(t), B=u(t-1), V; outputs C=u(t+1),
// D=u(t+2) -- out of place, so no thread ever reads a location another one writes in the launch.
// DRAM traffic: 3 reads + 2 writes per point per TWO levels = 10 B/point/level (vs 16 single level).
//
// Cross-check: bitwise against two launches of the single-level kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -o kbench2 kbench2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct Args {
  const float *A, *B, *V; float *C, *D;
  long long pitch; int ncol4, row0, row1, rows_per_cta;
  int lap_i0, lap_i1, lap_j0, lap_j1;
  int src_on, src_gi, src_j; float amp1, amp2;
  float cz[9], cx[9];
};
__device__ __forceinline__ float4 ld4(const float* p){ return *reinterpret_cast<const float4*>(p);}
__device__ __forceinline__ float4 ldnc(const float* p){ return __ldg(reinterpret_cast<const float4*>(p));}
__device__ __forceinline__ float getk(const float4& v,int k){ return k==0?v.x:k==1?v.y:k==2?v.z:v.w; }
__device__ __forceinline__ float leap(float p,float pp,float t){ double d=__fma_rn(2.0,(double)p,-(double)pp); return __double2float_rn(__dadd_rn(d,(double)t)); }
__device__ __forceinline__ float4 shup(const float4& v){ return make_float4(__shfl_up_sync(~0u,v.x,1),__shfl_up_sync(~0u,v.y,1),__shfl_up_sync(~0u,v.z,1),__shfl_up_sync(~0u,v.w,1)); }
__device__ __forceinline__ float4 shdn(const float4& v){ return make_float4(__shfl_down_sync(~0u,v.x,1),__shfl_down_sync(~0u,v.y,1),__shfl_down_sync(~0u,v.z,1),__shfl_down_sync(~0u,v.w,1)); }


// ---- recipe FAST (symmetric pairs + FMA, float update): one Laplacian + update for the 4 samples of a thread
template<int W> __device__ __forceinline__ void fast_row(const Args& a,const float* za,const float4* w,int u,const float4& c4,const float4& o,const float4& v,float* res,bool ring,int row,int j0){
  constexpr int H=4;
  #pragma unroll
  for(int k=0;k<4;k++){
    float s=getk(c4,k)*(a.cz[H]+a.cx[H]);
    #pragma unroll
    for(int d=1;d<=H;d++){
      s=__fmaf_rn(a.cz[H+d], za[4+k-d]+za[4+k+d], s);
      s=__fmaf_rn(a.cx[H+d], getk(w[(u+H-d)%W],k)+getk(w[(u+H+d)%W],k), s);
    }
    if(ring && (row<a.lap_i0||row>=a.lap_i1||j0+k<a.lap_j0||j0+k>=a.lap_j1)) s=0.f;
    res[k]=__fmaf_rn(getk(v,k), s, 2.0f*getk(c4,k)-getk(o,k));
  }
}

// ------------------------------------------------------------------ single level (production kernel's plain path)
template<int MINB>
__global__ void __launch_bounds__(256,MINB) k1(const __grid_constant__ Args a){
  constexpr int ORDER=8,H=4,W=9;
  const int q=blockIdx.x*blockDim.x+threadIdx.x; if(q>=a.ncol4) return;
  const int j0=q*4;
  const int rb=a.row0+blockIdx.y*a.rows_per_cta; const int re=min(rb+a.rows_per_cta,a.row1); if(rb>=re) return;
  const long long pitch=a.pitch;
  const bool ring = j0<a.lap_j0 || j0+4>a.lap_j1 || rb<a.lap_i0 || re>a.lap_i1;
  const bool near_src = a.src_on && a.src_j>=j0 && a.src_j<j0+4;
  const float* __restrict__ pc=a.A+j0+(long long)(rb-H)*pitch;
  float* __restrict__ ppc=a.C+j0+(long long)rb*pitch;       // in place: C holds u(t-1) in, u(t+1) out
  const float* __restrict__ vc=a.V+j0+(long long)rb*pitch;
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
          for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res[k2]=__fadd_rn(res[k2],a.amp1);
        }
        *reinterpret_cast<float4*>(ppc)=make_float4(res[0],res[1],res[2],res[3]);
        pc+=pitch; ppc+=pitch; vc+=pitch;
      }
    }
  }
}

template<int MINB>
__global__ void __launch_bounds__(256,MINB) k1f(const __grid_constant__ Args a){
  constexpr int ORDER=8,H=4,W=9;
  const int q=blockIdx.x*blockDim.x+threadIdx.x; if(q>=a.ncol4) return;
  const int j0=q*4;
  const int rb=a.row0+blockIdx.y*a.rows_per_cta; const int re=min(rb+a.rows_per_cta,a.row1); if(rb>=re) return;
  const long long pitch=a.pitch;
  const bool ring = j0<a.lap_j0 || j0+4>a.lap_j1 || rb<a.lap_i0 || re>a.lap_i1;
  const bool near_src = a.src_on && a.src_j>=j0 && a.src_j<j0+4;
  const float* __restrict__ pc=a.A+j0+(long long)(rb-H)*pitch;
  float* __restrict__ ppc=a.C+j0+(long long)rb*pitch;       // in place: C holds u(t-1) in, u(t+1) out
  const float* __restrict__ vc=a.V+j0+(long long)rb*pitch;
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
        float res[4]; fast_row<W>(a,za,w,u,c4,o,v,res,ring,re-left+u,j0);
        if(near_src && re-left+u==a.src_gi){
          #pragma unroll
          for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res[k2]=__fadd_rn(res[k2],a.amp1);
        }
        *reinterpret_cast<float4*>(ppc)=make_float4(res[0],res[1],res[2],res[3]);
        pc+=pitch; ppc+=pitch; vc+=pitch;
      }
    }
  }
}


// ------------------------------------------------------------------ two levels per pass
// PF bit0: register prefetch of the next row's operands; PF 2/3: prefetch.global.L2, 4/5: prefetch.global.L1, D rows ahead
template<int PF> __device__ __forceinline__ void pfx(const void* p){
  if(PF==2||PF==3) asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
  else asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
}
struct RowIn { float4 wn,l,r,o,v; };

template<int NT,int MINB,int PF,int D>
__global__ void __launch_bounds__(NT,MINB) k2(const __grid_constant__ Args a){
  constexpr int ORDER=8,H=4,W=9;
  const int lane=threadIdx.x&31;
  const int wg=blockIdx.x*(NT/32)+(threadIdx.x>>5);
  if(wg*30>=a.ncol4) return;                               // whole warp beyond the grid
  const int q=wg*30+lane-1;                                 // float4 column; -1 / >= ncol4 on clamped halo lanes
  const bool own = lane>=1 && lane<=30 && q<a.ncol4;        // this lane stores its column
  const int qc=min(max(q,0),a.ncol4-1);
  const int j0=q*4;
  const int rb=a.row0+blockIdx.y*a.rows_per_cta; const int re=min(rb+a.rows_per_cta,a.row1); if(rb>=re) return;
  const long long pitch=a.pitch;
  const bool ring = j0<a.lap_j0 || j0+4>a.lap_j1 || rb-H<a.lap_i0 || re+H>a.lap_i1;
  const bool near_src = a.src_on && a.src_j>=j0 && a.src_j<j0+4;
  const float* __restrict__ pA=a.A+4*qc+(long long)(rb-2*H)*pitch;     // row being streamed in (rA+H)
  const long long offA=4*qc+(long long)(rb-H)*pitch;                    // row rA
  const float* __restrict__ pB=a.B+offA;
  const float* __restrict__ pV=a.V+offA;
  float* __restrict__ pC=a.C+offA;
  float* __restrict__ pD=a.D+offA-(long long)H*pitch;                   // row rB = rA-H
  float4 w1[W], w2[W];
  #pragma unroll
  for(int s=0;s<2*H;s++){ w1[s]=ld4(pA); pA+=pitch; }
  #pragma unroll
  for(int s=0;s<W;s++) w2[s]=make_float4(0.f,0.f,0.f,0.f);
  auto loadrow=[&](const float* xa,const float* xb,const float* xv){ RowIn r; r.wn=ld4(xa); const float* ctr=xa-(long long)H*pitch; r.l=ld4(ctr-4); r.r=ld4(ctr+4); r.o=ld4(xb); r.v=ldnc(xv); return r; };
  RowIn cur; if(PF&1) cur=loadrow(pA,pB,pV);
  const bool pfl=(lane&7)==0;
  if(PF>=2 && pfl){
    #pragma unroll
    for(int d=0;d<D;d++){ pfx<PF>(pA+d*pitch); pfx<PF>(pB+d*pitch); pfx<PF>(pV+d*pitch); }
  }
  int rA=rb-H;
  for(int left=re-rb+2*H; left>0; left-=W){
    #pragma unroll
    for(int u=0;u<W;u++){
      if(u<left){
        // ---------------- level t -> t+1 at row rA
        if(PF>=2 && pfl){ pfx<PF>(pA+D*pitch); pfx<PF>(pB+D*pitch); pfx<PF>(pV+D*pitch); }
        if(!(PF&1)) cur=loadrow(pA,pB,pV);
        w1[(u+2*H)%W]=cur.wn;
        const float4 c4=w1[(u+H)%W];
        float res1[4];
        {
          const float za[12]={cur.l.x,cur.l.y,cur.l.z,cur.l.w,c4.x,c4.y,c4.z,c4.w,cur.r.x,cur.r.y,cur.r.z,cur.r.w};
          float lap[4];
          #pragma unroll
          for(int k2=0;k2<4;k2++){
            float az=__fmul_rn(za[k2],a.cz[0]); float ax=__fmul_rn(getk(w1[u%W],k2),a.cx[0]);
            #pragma unroll
            for(int io=1;io<=ORDER;io++){ az=__fadd_rn(az,__fmul_rn(za[k2+io],a.cz[io])); ax=__fadd_rn(ax,__fmul_rn(getk(w1[(u+io)%W],k2),a.cx[io])); }
            lap[k2]=__fadd_rn(az,ax);
          }
          if(ring){ const bool rin=rA>=a.lap_i0&&rA<a.lap_i1;
            #pragma unroll
            for(int k2=0;k2<4;k2++) if(!rin||j0+k2<a.lap_j0||j0+k2>=a.lap_j1) lap[k2]=0.f; }
          #pragma unroll
          for(int k2=0;k2<4;k2++) res1[k2]=leap(getk(c4,k2),getk(cur.o,k2),__fmul_rn(getk(cur.v,k2),lap[k2]));
          if(near_src && rA==a.src_gi){
            #pragma unroll
            for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res1[k2]=__fadd_rn(res1[k2],a.amp1);
          }
        }
        const float4 n1=make_float4(res1[0],res1[1],res1[2],res1[3]);
        if(own && rA>=rb && rA<re) *reinterpret_cast<float4*>(pC)=n1;
        w2[(u+2*H)%W]=n1;
        if(PF&1){ if(left-u>1) cur=loadrow(pA+pitch,pB+pitch,pV+pitch); }
        // ---------------- level t+1 -> t+2 at row rB = rA-4 (window complete once rA >= rb+4)
        if(rA>=rb+H){
          const int rB=rA-H;
          const float4 c2=w2[(u+H)%W];
          const float4 l2=shup(c2), r2=shdn(c2);
          const float4 vb=ldnc(pV-(long long)H*pitch);
          const float4 o2=w1[u%W];                           // u(t)[rB]
          const float za[12]={l2.x,l2.y,l2.z,l2.w,c2.x,c2.y,c2.z,c2.w,r2.x,r2.y,r2.z,r2.w};
          float lap[4],res2[4];
          #pragma unroll
          for(int k2=0;k2<4;k2++){
            float az=__fmul_rn(za[k2],a.cz[0]); float ax=__fmul_rn(getk(w2[u%W],k2),a.cx[0]);
            #pragma unroll
            for(int io=1;io<=ORDER;io++){ az=__fadd_rn(az,__fmul_rn(za[k2+io],a.cz[io])); ax=__fadd_rn(ax,__fmul_rn(getk(w2[(u+io)%W],k2),a.cx[io])); }
            lap[k2]=__fadd_rn(az,ax);
          }
          if(ring){ const bool rin=rB>=a.lap_i0&&rB<a.lap_i1;
            #pragma unroll
            for(int k2=0;k2<4;k2++) if(!rin||j0+k2<a.lap_j0||j0+k2>=a.lap_j1) lap[k2]=0.f; }
          #pragma unroll
          for(int k2=0;k2<4;k2++) res2[k2]=leap(getk(c2,k2),getk(o2,k2),__fmul_rn(getk(vb,k2),lap[k2]));
          if(near_src && rB==a.src_gi){
            #pragma unroll
            for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res2[k2]=__fadd_rn(res2[k2],a.amp2);
          }
          if(own) *reinterpret_cast<float4*>(pD)=make_float4(res2[0],res2[1],res2[2],res2[3]);
        }
        pA+=pitch; pB+=pitch; pV+=pitch; pC+=pitch; pD+=pitch; rA++;
      }
    }
  }
}


template<int NT,int MINB,int PF,int D>
__global__ void __launch_bounds__(NT,MINB) k2f(const __grid_constant__ Args a){
  constexpr int ORDER=8,H=4,W=9;
  const int lane=threadIdx.x&31;
  const int wg=blockIdx.x*(NT/32)+(threadIdx.x>>5);
  if(wg*30>=a.ncol4) return;                               // whole warp beyond the grid
  const int q=wg*30+lane-1;                                 // float4 column; -1 / >= ncol4 on clamped halo lanes
  const bool own = lane>=1 && lane<=30 && q<a.ncol4;        // this lane stores its column
  const int qc=min(max(q,0),a.ncol4-1);
  const int j0=q*4;
  const int rb=a.row0+blockIdx.y*a.rows_per_cta; const int re=min(rb+a.rows_per_cta,a.row1); if(rb>=re) return;
  const long long pitch=a.pitch;
  const bool ring = j0<a.lap_j0 || j0+4>a.lap_j1 || rb-H<a.lap_i0 || re+H>a.lap_i1;
  const bool near_src = a.src_on && a.src_j>=j0 && a.src_j<j0+4;
  const float* __restrict__ pA=a.A+4*qc+(long long)(rb-2*H)*pitch;     // row being streamed in (rA+H)
  const long long offA=4*qc+(long long)(rb-H)*pitch;                    // row rA
  const float* __restrict__ pB=a.B+offA;
  const float* __restrict__ pV=a.V+offA;
  float* __restrict__ pC=a.C+offA;
  float* __restrict__ pD=a.D+offA-(long long)H*pitch;                   // row rB = rA-H
  float4 w1[W], w2[W];
  #pragma unroll
  for(int s=0;s<2*H;s++){ w1[s]=ld4(pA); pA+=pitch; }
  #pragma unroll
  for(int s=0;s<W;s++) w2[s]=make_float4(0.f,0.f,0.f,0.f);
  auto loadrow=[&](const float* xa,const float* xb,const float* xv){ RowIn r; r.wn=ld4(xa); const float* ctr=xa-(long long)H*pitch; r.l=ld4(ctr-4); r.r=ld4(ctr+4); r.o=ld4(xb); r.v=ldnc(xv); return r; };
  RowIn cur; if(PF&1) cur=loadrow(pA,pB,pV);
  const bool pfl=(lane&7)==0;
  if(PF>=2 && pfl){
    #pragma unroll
    for(int d=0;d<D;d++){ pfx<PF>(pA+d*pitch); pfx<PF>(pB+d*pitch); pfx<PF>(pV+d*pitch); }
  }
  int rA=rb-H;
  for(int left=re-rb+2*H; left>0; left-=W){
    #pragma unroll
    for(int u=0;u<W;u++){
      if(u<left){
        // ---------------- level t -> t+1 at row rA
        if(PF>=2 && pfl){ pfx<PF>(pA+D*pitch); pfx<PF>(pB+D*pitch); pfx<PF>(pV+D*pitch); }
        if(!(PF&1)) cur=loadrow(pA,pB,pV);
        w1[(u+2*H)%W]=cur.wn;
        const float4 c4=w1[(u+H)%W];
        float res1[4];
        {
          const float za[12]={cur.l.x,cur.l.y,cur.l.z,cur.l.w,c4.x,c4.y,c4.z,c4.w,cur.r.x,cur.r.y,cur.r.z,cur.r.w};
          fast_row<W>(a,za,w1,u,c4,cur.o,cur.v,res1,ring,rA,j0);
          if(near_src && rA==a.src_gi){
            #pragma unroll
            for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res1[k2]=__fadd_rn(res1[k2],a.amp1);
          }
        }
        const float4 n1=make_float4(res1[0],res1[1],res1[2],res1[3]);
        if(own && rA>=rb && rA<re) *reinterpret_cast<float4*>(pC)=n1;
        w2[(u+2*H)%W]=n1;
        if(PF&1){ if(left-u>1) cur=loadrow(pA+pitch,pB+pitch,pV+pitch); }
        // ---------------- level t+1 -> t+2 at row rB = rA-4 (window complete once rA >= rb+4)
        if(rA>=rb+H){
          const int rB=rA-H;
          const float4 c2=w2[(u+H)%W];
          const float4 l2=shup(c2), r2=shdn(c2);
          const float4 vb=ldnc(pV-(long long)H*pitch);
          const float4 o2=w1[u%W];                           // u(t)[rB]
          const float za[12]={l2.x,l2.y,l2.z,l2.w,c2.x,c2.y,c2.z,c2.w,r2.x,r2.y,r2.z,r2.w};
          float res2[4]; fast_row<W>(a,za,w2,u,c2,o2,vb,res2,ring,rB,j0);
          if(near_src && rB==a.src_gi){
            #pragma unroll
            for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res2[k2]=__fadd_rn(res2[k2],a.amp2);
          }
          if(own) *reinterpret_cast<float4*>(pD)=make_float4(res2[0],res2[1],res2[2],res2[3]);
        }
        pA+=pitch; pB+=pitch; pV+=pitch; pC+=pitch; pD+=pitch; rA++;
      }
    }
  }
}



// ------------------------------------------------------------------ two levels per pass, TMA-staged inputs
// The CTA's rows of A, B, V are staged through 9-deep shared-memory rings by 1-D bulk copies
// (cp.async.bulk + mbarrier complete_tx), issued PD steps ahead by one elected thread; the warps stay
// independent in their arithmetic (halo lanes, shuffles) and only meet at the ring's full/empty mbarriers.
__device__ __forceinline__ unsigned smem_u32(const void* p){ return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar,unsigned cnt){ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar),"r"(cnt) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar,unsigned bytes){ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar),"r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned bar){ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar,unsigned parity){
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" :: "r"(bar),"r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst,const void* src,unsigned bytes,unsigned bar){
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst),"l"(src),"r"(bytes),"r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds4(unsigned addr){ float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w) : "r"(addr)); return v; }

template<int NW,int MINB,int PD>
__global__ void __launch_bounds__(NW*32,MINB) k2t(const __grid_constant__ Args a){
  constexpr int ORDER=8,H=4,W=9,NS=9;
  constexpr unsigned SEG=(NW*30+4)*16;                      // bytes per staged row: 30 float4 per warp + 2 halo float4 each side
  extern __shared__ __align__(128) unsigned char smem[];
  const unsigned sA=smem_u32(smem), sB=sA+NS*SEG, sV=sB+NS*SEG, sFull=sV+NS*SEG, sEmpty=sFull+NS*8;
  const int lane=threadIdx.x&31, warp=threadIdx.x>>5;
  const int Q0=blockIdx.x*(NW*30);                          // first owned float4 column of the CTA
  const int nwa=min(NW,(a.ncol4-Q0+29)/30);                 // warps with work
  const int rb=a.row0+blockIdx.y*a.rows_per_cta; const int re=min(rb+a.rows_per_cta,a.row1);
  if(rb>=re) return;
  if(threadIdx.x==0){
    for(int s=0;s<NS;s++){ mbar_init(sFull+8*s,1); mbar_init(sEmpty+8*s,nwa); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if(warp>=nwa) return;
  const int q=Q0+warp*30+lane-1;
  const bool own = lane>=1 && lane<=30 && q<a.ncol4;
  const int j0=q*4;
  const long long pitch=a.pitch;
  const bool ring = j0<a.lap_j0 || j0+4>a.lap_j1 || rb-H<a.lap_i0 || re+H>a.lap_i1;
  const bool near_src = a.src_on && a.src_j>=j0 && a.src_j<j0+4;
  const unsigned toff=(warp*30+lane+1)*16;                  // this thread's float4 inside a staged row
  const int T=re-rb+4*H;                                    // steps: A rows rb-8 .. re+7
  const bool producer = threadIdx.x==0;
  // bundle of step tau: A[rb-8+tau]; and B,V[rb-12+tau] once tau >= 8 (first row any stage A needs)
  const long long g0=4*(long long)(Q0-2)+(long long)(rb-2*H)*pitch;
  auto issue=[&](int tau,int slot){
    const unsigned bar=sFull+8*slot; const bool full=tau>=2*H;
    mbar_expect_tx(bar, full?3*SEG:SEG);
    const long long g=g0+(long long)tau*pitch;
    bulk_g2s(sA+slot*SEG, a.A+g, SEG, bar);
    if(full){ bulk_g2s(sB+slot*SEG, a.B+g-H*pitch, SEG, bar); bulk_g2s(sV+slot*SEG, a.V+g-H*pitch, SEG, bar); }
  };
  if(producer){
    #pragma unroll
    for(int t0=0;t0<PD;t0++) if(t0<T) issue(t0,t0);
  }
  long long off=4*(long long)q+(long long)(rb-3*H)*pitch;   // row rA of step 0 (= rb-12; stores start later)
  float4 w1[W], w2[W];
  #pragma unroll
  for(int s=0;s<W;s++){ w1[s]=make_float4(0.f,0.f,0.f,0.f); w2[s]=w1[s]; }
  unsigned ph=0; int t=0;
  for(int left=T; left>0; left-=W, ph^=1u){
    #pragma unroll
    for(int u=0;u<W;u++){
      if(u<left){
        if(producer && t+PD<T){
          // the slot of step t+PD was last read at step t+PD-9+4: wait until every warp has finished it
          const int tw=t+PD-5;
          if(tw>=0){ const int uw=(u+PD+4)%NS; mbar_wait(sEmpty+8*uw, ((unsigned)(tw/NS))&1u); }
          issue(t+PD,(u+PD)%NS);
        }
        mbar_wait(sFull+8*u, ph);
        const int rA=rb-3*H+t;                               // = r-4, r = rb-8+t
        w1[(u+2*H)%W]=lds4(sA+u*SEG+toff);
        if(t>=2*H){
          // ---------------- level t -> t+1 at row rA
          const unsigned ctr=sA+((u+5)%NS)*SEG+toff;
          const float4 l=lds4(ctr-16), r=lds4(ctr+16), o=lds4(sB+u*SEG+toff), v=lds4(sV+u*SEG+toff);
          const float4 c4=w1[(u+H)%W];
          const float za[12]={l.x,l.y,l.z,l.w,c4.x,c4.y,c4.z,c4.w,r.x,r.y,r.z,r.w};
          float lap[4],res1[4];
          #pragma unroll
          for(int k2=0;k2<4;k2++){
            float az=__fmul_rn(za[k2],a.cz[0]); float ax=__fmul_rn(getk(w1[u%W],k2),a.cx[0]);
            #pragma unroll
            for(int io=1;io<=ORDER;io++){ az=__fadd_rn(az,__fmul_rn(za[k2+io],a.cz[io])); ax=__fadd_rn(ax,__fmul_rn(getk(w1[(u+io)%W],k2),a.cx[io])); }
            lap[k2]=__fadd_rn(az,ax);
          }
          if(ring){ const bool rin=rA>=a.lap_i0&&rA<a.lap_i1;
            #pragma unroll
            for(int k2=0;k2<4;k2++) if(!rin||j0+k2<a.lap_j0||j0+k2>=a.lap_j1) lap[k2]=0.f; }
          #pragma unroll
          for(int k2=0;k2<4;k2++) res1[k2]=leap(getk(c4,k2),getk(o,k2),__fmul_rn(getk(v,k2),lap[k2]));
          if(near_src && rA==a.src_gi){
            #pragma unroll
            for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res1[k2]=__fadd_rn(res1[k2],a.amp1);
          }
          const float4 n1=make_float4(res1[0],res1[1],res1[2],res1[3]);
          if(own && rA>=rb && rA<re) *reinterpret_cast<float4*>(a.C+off)=n1;
          w2[(u+2*H)%W]=n1;
        }
        if(t>=4*H){
          // ---------------- level t+1 -> t+2 at row rB = rA-4
          const int rB=rA-H;
          const float4 c2=w2[(u+H)%W];
          const float4 l2=shup(c2), r2=shdn(c2);
          const float4 vb=lds4(sV+((u+5)%NS)*SEG+toff);
          const float4 o2=w1[u%W];
          const float za[12]={l2.x,l2.y,l2.z,l2.w,c2.x,c2.y,c2.z,c2.w,r2.x,r2.y,r2.z,r2.w};
          float lap[4],res2[4];
          #pragma unroll
          for(int k2=0;k2<4;k2++){
            float az=__fmul_rn(za[k2],a.cz[0]); float ax=__fmul_rn(getk(w2[u%W],k2),a.cx[0]);
            #pragma unroll
            for(int io=1;io<=ORDER;io++){ az=__fadd_rn(az,__fmul_rn(za[k2+io],a.cz[io])); ax=__fadd_rn(ax,__fmul_rn(getk(w2[(u+io)%W],k2),a.cx[io])); }
            lap[k2]=__fadd_rn(az,ax);
          }
          if(ring){ const bool rin=rB>=a.lap_i0&&rB<a.lap_i1;
            #pragma unroll
            for(int k2=0;k2<4;k2++) if(!rin||j0+k2<a.lap_j0||j0+k2>=a.lap_j1) lap[k2]=0.f; }
          #pragma unroll
          for(int k2=0;k2<4;k2++) res2[k2]=leap(getk(c2,k2),getk(o2,k2),__fmul_rn(getk(vb,k2),lap[k2]));
          if(near_src && rB==a.src_gi){
            #pragma unroll
            for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res2[k2]=__fadd_rn(res2[k2],a.amp2);
          }
          if(own) *reinterpret_cast<float4*>(a.D+off-H*pitch)=make_float4(res2[0],res2[1],res2[2],res2[3]);
        }
        __syncwarp();
        if(lane==0) mbar_arrive(sEmpty+8*u);
        off+=pitch; t++;
      }
    }
  }
}

// ------------------------------------------------------------------ k2u: k2t with a branch-free steady-state body
// Every step runs both stages unconditionally (the first 16 steps of a chunk compute on not-yet-valid
// windows and simply do not store), the step count is padded to a multiple of 9, and the only branches
// left are the rare ring/source path and the producer's issue block.
template<int NW,int MINB,int PD,int SYM>
__global__ void __launch_bounds__(NW*32,MINB) k2u(const __grid_constant__ Args a){
  constexpr int ORDER=8,H=4,W=9,NS=9;
  constexpr unsigned SEG=(NW*30+4)*16;
  extern __shared__ __align__(128) unsigned char smem[];
  const unsigned sA=smem_u32(smem), sB=sA+NS*SEG, sV=sB+NS*SEG, sFull=sV+NS*SEG, sEmpty=sFull+NS*8;
  const int lane=threadIdx.x&31, warp=threadIdx.x>>5;
  const int Q0=blockIdx.x*(NW*30);
  const int nwa=min(NW,(a.ncol4-Q0+29)/30);
  const int rb=a.row0+blockIdx.y*a.rows_per_cta; const int re=min(rb+a.rows_per_cta,a.row1);
  if(rb>=re) return;
  if(threadIdx.x==0){
    for(int s=0;s<NS;s++){ mbar_init(sFull+8*s,1); mbar_init(sEmpty+8*s,nwa); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if(warp>=nwa) return;
  const int q=Q0+warp*30+lane-1;
  const bool own = lane>=1 && lane<=30 && q<a.ncol4;
  const int j0=q*4;
  const long long pitch=a.pitch;
  const bool rare = (j0<a.lap_j0 || j0+4>a.lap_j1 || rb-H<a.lap_i0 || re+H>a.lap_i1) || (a.src_on && a.src_j>=j0 && a.src_j<j0+4);
  const unsigned toff=(warp*30+lane+1)*16;
  const int T=((re-rb+4*H+W-1)/W)*W;                        // steps, padded to whole unrolled blocks
  const long long g0=4*(long long)(Q0-2)+(long long)(rb-2*H)*pitch;
  auto issue=[&](int tau,int slot){
    const unsigned bar=sFull+8*slot;
    mbar_expect_tx(bar,3*SEG);
    const long long g=g0+(long long)tau*pitch;
    bulk_g2s(sA+slot*SEG, a.A+g, SEG, bar);
    bulk_g2s(sB+slot*SEG, a.B+g-H*pitch, SEG, bar);
    bulk_g2s(sV+slot*SEG, a.V+g-H*pitch, SEG, bar);
  };
  if(threadIdx.x==0){
    #pragma unroll
    for(int t0=0;t0<PD;t0++) issue(t0,t0);
  }
  int pw=0;                                                 // warp whose lane 0 issues this step's bundle (round robin)
  float* __restrict__ pC=a.C+4*(long long)q+(long long)(rb-3*H)*pitch;   // row rA of step 0
  float* __restrict__ pD=a.D+4*(long long)q+(long long)(rb-4*H)*pitch;   // row rB of step 0
  float4 w1[W], w2[W];
  #pragma unroll
  for(int s=0;s<W;s++){ w1[s]=make_float4(0.f,0.f,0.f,0.f); w2[s]=w1[s]; }
  unsigned ph=0; int rA=rb-3*H;
  for(int t=0; t<T; t+=W, ph^=1u){
    #pragma unroll
    for(int u=0;u<W;u++){
      if(warp==pw && lane==0 && t+u+PD<T){
        // slot (u+PD)%9 was last read at step t+u+PD-5 (static parity: that step lies in this or the previous block)
        if(u+PD-5>=0) mbar_wait(sEmpty+8*((u+PD+4)%NS), ph);
        else if(t>0) mbar_wait(sEmpty+8*((u+PD+4)%NS), ph^1u);
        issue(t+u+PD,(u+PD)%NS);
      }
      pw=(pw+1==nwa)?0:pw+1;
      mbar_wait(sFull+8*u, ph);
      // ---------------- level t -> t+1 at row rA
      w1[(u+2*H)%W]=lds4(sA+u*SEG+toff);
      const unsigned ctr=sA+((u+5)%NS)*SEG+toff;
      const float4 l=lds4(ctr-16), r=lds4(ctr+16), o=lds4(sB+u*SEG+toff), v=lds4(sV+u*SEG+toff);
      const float4 vb=lds4(sV+((u+5)%NS)*SEG+toff);
      const float4 c4=w1[(u+H)%W];
      float lap[4],res1[4];
      {
        const float za[12]={l.x,l.y,l.z,l.w,c4.x,c4.y,c4.z,c4.w,r.x,r.y,r.z,r.w};
        #pragma unroll
        for(int k2=0;k2<4;k2++){
          float az=__fmul_rn(za[k2],a.cz[0]); float ax=__fmul_rn(getk(w1[u%W],k2),a.cx[0]);
          #pragma unroll
          for(int io=1;io<=ORDER;io++){ az=__fadd_rn(az,__fmul_rn(za[k2+io],a.cz[SYM&&io>H?ORDER-io:io])); ax=__fadd_rn(ax,__fmul_rn(getk(w1[(u+io)%W],k2),a.cx[SYM&&io>H?ORDER-io:io])); }
          lap[k2]=__fadd_rn(az,ax);
        }
      }
      if(rare){ const bool rin=rA>=a.lap_i0&&rA<a.lap_i1;
        #pragma unroll
        for(int k2=0;k2<4;k2++) if(!rin||j0+k2<a.lap_j0||j0+k2>=a.lap_j1) lap[k2]=0.f; }
      #pragma unroll
      for(int k2=0;k2<4;k2++) res1[k2]=leap(getk(c4,k2),getk(o,k2),__fmul_rn(getk(v,k2),lap[k2]));
      if(rare && a.src_on && rA==a.src_gi){
        #pragma unroll
        for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res1[k2]=__fadd_rn(res1[k2],a.amp1);
      }
      const float4 n1=make_float4(res1[0],res1[1],res1[2],res1[3]);
      if(own && rA>=rb && rA<re) *reinterpret_cast<float4*>(pC)=n1;
      w2[(u+2*H)%W]=n1;
      // ---------------- level t+1 -> t+2 at row rB = rA-4
      {
        const int rB=rA-H;
        const float4 c2=w2[(u+H)%W];
        const float4 l2=shup(c2), r2=shdn(c2);
        const float4 o2=w1[u%W];
        const float za[12]={l2.x,l2.y,l2.z,l2.w,c2.x,c2.y,c2.z,c2.w,r2.x,r2.y,r2.z,r2.w};
        float res2[4];
        #pragma unroll
        for(int k2=0;k2<4;k2++){
          float az=__fmul_rn(za[k2],a.cz[0]); float ax=__fmul_rn(getk(w2[u%W],k2),a.cx[0]);
          #pragma unroll
          for(int io=1;io<=ORDER;io++){ az=__fadd_rn(az,__fmul_rn(za[k2+io],a.cz[SYM&&io>H?ORDER-io:io])); ax=__fadd_rn(ax,__fmul_rn(getk(w2[(u+io)%W],k2),a.cx[SYM&&io>H?ORDER-io:io])); }
          lap[k2]=__fadd_rn(az,ax);
        }
        if(rare){ const bool rin=rB>=a.lap_i0&&rB<a.lap_i1;
          #pragma unroll
          for(int k2=0;k2<4;k2++) if(!rin||j0+k2<a.lap_j0||j0+k2>=a.lap_j1) lap[k2]=0.f; }
        #pragma unroll
        for(int k2=0;k2<4;k2++) res2[k2]=leap(getk(c2,k2),getk(o2,k2),__fmul_rn(getk(vb,k2),lap[k2]));
        if(rare && a.src_on && rB==a.src_gi){
          #pragma unroll
          for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res2[k2]=__fadd_rn(res2[k2],a.amp2);
        }
        if(own && rB>=rb && rB<re) *reinterpret_cast<float4*>(pD)=make_float4(res2[0],res2[1],res2[2],res2[3]);
      }
      __syncwarp();
      if(lane==0) mbar_arrive(sEmpty+8*u);
      pC+=pitch; pD+=pitch; rA++;
    }
  }
}

// ------------------------------------------------------------------ k2w: per-WARP private TMA rings (no coupling between warps)
// Every warp stages its own 34 (A) / 32 (B, V) float4 columns; its lane 0 issues the bulk copies PD steps
// ahead (A,V: 9-deep rings, rows live 5 steps; B: 3-deep ring, 2 steps ahead).  A warp only ever waits on its
// own full barriers; slot reuse is ordered by the warp's own program order (+ __syncwarp).
template<int NW,int MINB,int PD,int SYM>
__global__ void __launch_bounds__(NW*32,MINB) k2w(const __grid_constant__ Args a){
  constexpr int ORDER=8,H=4,W=9,NS=9,NB=3,PB=2;
  constexpr unsigned SEGA=34*16, SEGV=32*16;
  constexpr unsigned WSM=NS*SEGA+NS*SEGV+NB*SEGV+(NS+NB)*8;   // bytes of shared memory per warp
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane=threadIdx.x&31, warp=threadIdx.x>>5;
  const unsigned sA=smem_u32(smem)+warp*WSM, sV=sA+NS*SEGA, sB=sV+NS*SEGV, sFull=sB+NB*SEGV, sFullB=sFull+NS*8;
  const int wg=blockIdx.x*NW+warp;
  const int rb=a.row0+blockIdx.y*a.rows_per_cta; const int re=min(rb+a.rows_per_cta,a.row1);
  if(rb>=re || wg*30>=a.ncol4) return;
  if(lane==0){
    for(int s=0;s<NS+NB;s++) mbar_init(sFull+8*s,1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  const int q=wg*30+lane-1;
  const bool own = lane>=1 && lane<=30 && q<a.ncol4;
  const int j0=q*4;
  const long long pitch=a.pitch;
  const bool rare = (j0<a.lap_j0 || j0+4>a.lap_j1 || rb-H<a.lap_i0 || re+H>a.lap_i1) || (a.src_on && a.src_j>=j0 && a.src_j<j0+4);
  const unsigned toffA=(lane+1)*16, toffV=lane*16;
  const int T=((re-rb+4*H+W-1)/W)*W;
  // global sources: A row rb-8+tau from column 4*(wg*30-2); B,V row rb-12+tau from column 4*(wg*30-1)
  const float* gA=a.A+4*(long long)(wg*30-2)+(long long)(rb-2*H)*pitch;
  const float* gV=a.V+4*(long long)(wg*30-1)+(long long)(rb-3*H)*pitch;
  const float* gB=a.B+4*(long long)(wg*30-1)+(long long)(rb-3*H)*pitch;
  if(lane==0){
    #pragma unroll
    for(int t0=0;t0<PD;t0++){
      mbar_expect_tx(sFull+8*t0,SEGA+SEGV);
      bulk_g2s(sA+t0*SEGA,gA+t0*pitch,SEGA,sFull+8*t0); bulk_g2s(sV+t0*SEGV,gV+t0*pitch,SEGV,sFull+8*t0);
    }
    #pragma unroll
    for(int t0=0;t0<PB;t0++){ mbar_expect_tx(sFullB+8*t0,SEGV); bulk_g2s(sB+t0*SEGV,gB+t0*pitch,SEGV,sFullB+8*t0); }
  }
  gA+=PD*pitch; gV+=PD*pitch; gB+=PB*pitch;                 // next rows to request
  float* __restrict__ pC=a.C+4*(long long)q+(long long)(rb-3*H)*pitch;
  float* __restrict__ pD=a.D+4*(long long)q+(long long)(rb-4*H)*pitch;
  float4 w1[W], w2[W];
  #pragma unroll
  for(int s=0;s<W;s++){ w1[s]=make_float4(0.f,0.f,0.f,0.f); w2[s]=w1[s]; }
  unsigned ph=0; int rA=rb-3*H;
  for(int t=0; t<T; t+=W, ph^=1u){
    #pragma unroll
    for(int u=0;u<W;u++){
      if(lane==0){
        if(t+u+PD<T){
          const unsigned bar=sFull+8*((u+PD)%NS);
          mbar_expect_tx(bar,SEGA+SEGV);
          bulk_g2s(sA+((u+PD)%NS)*SEGA,gA,SEGA,bar); bulk_g2s(sV+((u+PD)%NS)*SEGV,gV,SEGV,bar);
        }
        if(t+u+PB<T){
          const unsigned bar=sFullB+8*((u+PB)%NB);
          mbar_expect_tx(bar,SEGV);
          bulk_g2s(sB+((u+PB)%NB)*SEGV,gB,SEGV,bar);
        }
      }
      gA+=pitch; gV+=pitch; gB+=pitch;
      mbar_wait(sFull+8*u, ph);
      // ---------------- level t -> t+1 at row rA
      w1[(u+2*H)%W]=lds4(sA+u*SEGA+toffA);
      const unsigned ctr=sA+((u+5)%NS)*SEGA+toffA;
      const float4 l=lds4(ctr-16), r=lds4(ctr+16), v=lds4(sV+u*SEGV+toffV);
      const float4 vb=lds4(sV+((u+5)%NS)*SEGV+toffV);
      mbar_wait(sFullB+8*(u%NB), ph^((u/NB)&1u));
      const float4 o=lds4(sB+(u%NB)*SEGV+toffV);
      const float4 c4=w1[(u+H)%W];
      float lap[4],res1[4];
      {
        const float za[12]={l.x,l.y,l.z,l.w,c4.x,c4.y,c4.z,c4.w,r.x,r.y,r.z,r.w};
        #pragma unroll
        for(int k2=0;k2<4;k2++){
          float az=__fmul_rn(za[k2],a.cz[0]); float ax=__fmul_rn(getk(w1[u%W],k2),a.cx[0]);
          #pragma unroll
          for(int io=1;io<=ORDER;io++){ az=__fadd_rn(az,__fmul_rn(za[k2+io],a.cz[SYM&&io>H?ORDER-io:io])); ax=__fadd_rn(ax,__fmul_rn(getk(w1[(u+io)%W],k2),a.cx[SYM&&io>H?ORDER-io:io])); }
          lap[k2]=__fadd_rn(az,ax);
        }
      }
      if(rare){ const bool rin=rA>=a.lap_i0&&rA<a.lap_i1;
        #pragma unroll
        for(int k2=0;k2<4;k2++) if(!rin||j0+k2<a.lap_j0||j0+k2>=a.lap_j1) lap[k2]=0.f; }
      #pragma unroll
      for(int k2=0;k2<4;k2++) res1[k2]=leap(getk(c4,k2),getk(o,k2),__fmul_rn(getk(v,k2),lap[k2]));
      if(rare && a.src_on && rA==a.src_gi){
        #pragma unroll
        for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res1[k2]=__fadd_rn(res1[k2],a.amp1);
      }
      const float4 n1=make_float4(res1[0],res1[1],res1[2],res1[3]);
      if(own && rA>=rb && rA<re) *reinterpret_cast<float4*>(pC)=n1;
      w2[(u+2*H)%W]=n1;
      // ---------------- level t+1 -> t+2 at row rB = rA-4
      {
        const int rB=rA-H;
        const float4 c2=w2[(u+H)%W];
        const float4 l2=shup(c2), r2=shdn(c2);
        const float4 o2=w1[u%W];
        const float za[12]={l2.x,l2.y,l2.z,l2.w,c2.x,c2.y,c2.z,c2.w,r2.x,r2.y,r2.z,r2.w};
        float res2[4];
        #pragma unroll
        for(int k2=0;k2<4;k2++){
          float az=__fmul_rn(za[k2],a.cz[0]); float ax=__fmul_rn(getk(w2[u%W],k2),a.cx[0]);
          #pragma unroll
          for(int io=1;io<=ORDER;io++){ az=__fadd_rn(az,__fmul_rn(za[k2+io],a.cz[SYM&&io>H?ORDER-io:io])); ax=__fadd_rn(ax,__fmul_rn(getk(w2[(u+io)%W],k2),a.cx[SYM&&io>H?ORDER-io:io])); }
          lap[k2]=__fadd_rn(az,ax);
        }
        if(rare){ const bool rin=rB>=a.lap_i0&&rB<a.lap_i1;
          #pragma unroll
          for(int k2=0;k2<4;k2++) if(!rin||j0+k2<a.lap_j0||j0+k2>=a.lap_j1) lap[k2]=0.f; }
        #pragma unroll
        for(int k2=0;k2<4;k2++) res2[k2]=leap(getk(c2,k2),getk(o2,k2),__fmul_rn(getk(vb,k2),lap[k2]));
        if(rare && a.src_on && rB==a.src_gi){
          #pragma unroll
          for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res2[k2]=__fadd_rn(res2[k2],a.amp2);
        }
        if(own && rB>=rb && rB<re) *reinterpret_cast<float4*>(pD)=make_float4(res2[0],res2[1],res2[2],res2[3]);
      }
      __syncwarp();
      pC+=pitch; pD+=pitch; rA++;
    }
  }
}
template<int NW,int MINB,int PD,int SYM>
__global__ void __launch_bounds__(NW*32,MINB) k2wf(const __grid_constant__ Args a){
  constexpr int ORDER=8,H=4,W=9,NS=9,NB=3,PB=2;
  constexpr unsigned SEGA=34*16, SEGV=32*16;
  constexpr unsigned WSM=NS*SEGA+NS*SEGV+NB*SEGV+(NS+NB)*8;   // bytes of shared memory per warp
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane=threadIdx.x&31, warp=threadIdx.x>>5;
  const unsigned sA=smem_u32(smem)+warp*WSM, sV=sA+NS*SEGA, sB=sV+NS*SEGV, sFull=sB+NB*SEGV, sFullB=sFull+NS*8;
  const int wg=blockIdx.x*NW+warp;
  const int rb=a.row0+blockIdx.y*a.rows_per_cta; const int re=min(rb+a.rows_per_cta,a.row1);
  if(rb>=re || wg*30>=a.ncol4) return;
  if(lane==0){
    for(int s=0;s<NS+NB;s++) mbar_init(sFull+8*s,1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  const int q=wg*30+lane-1;
  const bool own = lane>=1 && lane<=30 && q<a.ncol4;
  const int j0=q*4;
  const long long pitch=a.pitch;
  const bool rare = (j0<a.lap_j0 || j0+4>a.lap_j1 || rb-H<a.lap_i0 || re+H>a.lap_i1) || (a.src_on && a.src_j>=j0 && a.src_j<j0+4);
  const unsigned toffA=(lane+1)*16, toffV=lane*16;
  const int T=((re-rb+4*H+W-1)/W)*W;
  // global sources: A row rb-8+tau from column 4*(wg*30-2); B,V row rb-12+tau from column 4*(wg*30-1)
  const float* gA=a.A+4*(long long)(wg*30-2)+(long long)(rb-2*H)*pitch;
  const float* gV=a.V+4*(long long)(wg*30-1)+(long long)(rb-3*H)*pitch;
  const float* gB=a.B+4*(long long)(wg*30-1)+(long long)(rb-3*H)*pitch;
  if(lane==0){
    #pragma unroll
    for(int t0=0;t0<PD;t0++){
      mbar_expect_tx(sFull+8*t0,SEGA+SEGV);
      bulk_g2s(sA+t0*SEGA,gA+t0*pitch,SEGA,sFull+8*t0); bulk_g2s(sV+t0*SEGV,gV+t0*pitch,SEGV,sFull+8*t0);
    }
    #pragma unroll
    for(int t0=0;t0<PB;t0++){ mbar_expect_tx(sFullB+8*t0,SEGV); bulk_g2s(sB+t0*SEGV,gB+t0*pitch,SEGV,sFullB+8*t0); }
  }
  gA+=PD*pitch; gV+=PD*pitch; gB+=PB*pitch;                 // next rows to request
  float* __restrict__ pC=a.C+4*(long long)q+(long long)(rb-3*H)*pitch;
  float* __restrict__ pD=a.D+4*(long long)q+(long long)(rb-4*H)*pitch;
  float4 w1[W], w2[W];
  #pragma unroll
  for(int s=0;s<W;s++){ w1[s]=make_float4(0.f,0.f,0.f,0.f); w2[s]=w1[s]; }
  unsigned ph=0; int rA=rb-3*H;
  for(int t=0; t<T; t+=W, ph^=1u){
    #pragma unroll
    for(int u=0;u<W;u++){
      if(lane==0){
        if(t+u+PD<T){
          const unsigned bar=sFull+8*((u+PD)%NS);
          mbar_expect_tx(bar,SEGA+SEGV);
          bulk_g2s(sA+((u+PD)%NS)*SEGA,gA,SEGA,bar); bulk_g2s(sV+((u+PD)%NS)*SEGV,gV,SEGV,bar);
        }
        if(t+u+PB<T){
          const unsigned bar=sFullB+8*((u+PB)%NB);
          mbar_expect_tx(bar,SEGV);
          bulk_g2s(sB+((u+PB)%NB)*SEGV,gB,SEGV,bar);
        }
      }
      gA+=pitch; gV+=pitch; gB+=pitch;
      mbar_wait(sFull+8*u, ph);
      // ---------------- level t -> t+1 at row rA
      w1[(u+2*H)%W]=lds4(sA+u*SEGA+toffA);
      const unsigned ctr=sA+((u+5)%NS)*SEGA+toffA;
      const float4 l=lds4(ctr-16), r=lds4(ctr+16), v=lds4(sV+u*SEGV+toffV);
      const float4 vb=lds4(sV+((u+5)%NS)*SEGV+toffV);
      mbar_wait(sFullB+8*(u%NB), ph^((u/NB)&1u));
      const float4 o=lds4(sB+(u%NB)*SEGV+toffV);
      const float4 c4=w1[(u+H)%W];
      float res1[4];
      {
        const float za[12]={l.x,l.y,l.z,l.w,c4.x,c4.y,c4.z,c4.w,r.x,r.y,r.z,r.w};
        fast_row<W>(a,za,w1,u,c4,o,v,res1,rare,rA,j0);
      }
      if(rare && a.src_on && rA==a.src_gi){
        #pragma unroll
        for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res1[k2]=__fadd_rn(res1[k2],a.amp1);
      }
      const float4 n1=make_float4(res1[0],res1[1],res1[2],res1[3]);
      if(own && rA>=rb && rA<re) *reinterpret_cast<float4*>(pC)=n1;
      w2[(u+2*H)%W]=n1;
      // ---------------- level t+1 -> t+2 at row rB = rA-4
      {
        const int rB=rA-H;
        const float4 c2=w2[(u+H)%W];
        const float4 l2=shup(c2), r2=shdn(c2);
        const float4 o2=w1[u%W];
        const float za[12]={l2.x,l2.y,l2.z,l2.w,c2.x,c2.y,c2.z,c2.w,r2.x,r2.y,r2.z,r2.w};
        float res2[4]; fast_row<W>(a,za,w2,u,c2,o2,vb,res2,rare,rB,j0);
        if(rare && a.src_on && rB==a.src_gi){
          #pragma unroll
          for(int k2=0;k2<4;k2++) if(j0+k2==a.src_j) res2[k2]=__fadd_rn(res2[k2],a.amp2);
        }
        if(own && rB>=rb && rB<re) *reinterpret_cast<float4*>(pD)=make_float4(res2[0],res2[1],res2[2],res2[3]);
      }
      __syncwarp();
      pC+=pitch; pD+=pitch; rA++;
    }
  }
}

template<int NW> constexpr size_t k2w_smem(){ return (size_t)NW*(9*544+9*512+3*512+12*8); }

template<int NW> constexpr size_t k2t_smem(){ return 3*9*(size_t)((NW*30+4)*16)+2*9*8; }

typedef void (*kfn)(const Args);
struct Var { const char* name; kfn f; int nt; size_t smem; };
#define VT(NW,MINB,PD) {"K2T_W" #NW "_B" #MINB "_P" #PD, k2t<NW,MINB,PD>, NW*32, k2t_smem<NW>()}
#define VU(NW,MINB,PD,SYM) {"K2U_W" #NW "_B" #MINB "_P" #PD "_S" #SYM, k2u<NW,MINB,PD,SYM>, NW*32, k2t_smem<NW>()}
#define VW(NW,MINB,PD,SYM) {"K2W_W" #NW "_B" #MINB "_P" #PD "_S" #SYM, k2w<NW,MINB,PD,SYM>, NW*32, k2w_smem<NW>()}
#define VWF(NW,MINB,PD) {"K2WF_W" #NW "_B" #MINB "_P" #PD, k2wf<NW,MINB,PD,0>, NW*32, k2w_smem<NW>()}
#define V2F(NT,MINB,PF,D) {"K2F_T" #NT "_B" #MINB "_PF" #PF "_D" #D, k2f<NT,MINB,PF,D>, NT, 0}
#define V2(NT,MINB,PF,D) {"K2_T" #NT "_B" #MINB "_PF" #PF "_D" #D, k2<NT,MINB,PF,D>, NT, 0}

static int cmp(const float* d1,const float* d2,size_t elems,std::vector<float>&x,std::vector<float>&y){
  cudaMemcpy(x.data(),d1,elems*4,cudaMemcpyDeviceToHost); cudaMemcpy(y.data(),d2,elems*4,cudaMemcpyDeviceToHost);
  return !memcmp(x.data(),y.data(),elems*4);
}

int main(int argc,char**argv){
  int n = argc>1?atoi(argv[1]):16384;
  const int G=16;                                  // guard rows above and below
  long long pitch=((long long)n+4+31)/32*32; size_t rows=n+2*G; size_t elems=rows*pitch;
  float *A,*B,*V,*C,*D,*R1,*R2;
  cudaMalloc(&A,elems*4); cudaMalloc(&B,elems*4); cudaMalloc(&V,elems*4); cudaMalloc(&C,elems*4); cudaMalloc(&D,elems*4);
  cudaMalloc(&R1,elems*4); cudaMalloc(&R2,elems*4);
  std::vector<float> h(elems,0.f),x(elems),y(elems);
  srand(1);
  auto fill=[&](float* d,int kind){ std::fill(h.begin(),h.end(),0.f);
    for(int i=0;i<n;i++) for(int j=0;j<n;j++) h[(size_t)(i+G)*pitch+j]= kind==2 ? 0.0625f+0.01f*((i*131+j)%7) : (rand()/(float)RAND_MAX-0.5f);
    cudaMemcpy(d,h.data(),elems*4,cudaMemcpyHostToDevice); };
  fill(A,0); fill(B,1); fill(V,2);
  Args a; memset(&a,0,sizeof a); a.pitch=pitch; a.ncol4=(n+3)/4; a.row0=0;a.row1=n;
  a.lap_i0=4;a.lap_i1=n-4;a.lap_j0=4;a.lap_j1=n-4; a.src_on=1;a.src_gi=n/2;a.src_j=41;a.amp1=0.5f;a.amp2=-0.25f;
  for(int i=0;i<9;i++){int m=i<=4?i:8-i; a.cz[i]=0.01f*(m+1)*(m%2?1:-1); a.cx[i]=0.02f*(9-m)*(m%2?-1:1);}
  const long long o0=(long long)G*pitch;
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  // ---- reference: two single-level launches.  R1 = u(t+1) (in place over a copy of B), R2 = u(t+2) (over a copy of A)
  {
    dim3 grid((a.ncol4+255)/256,(n+31)/32), block(256); Args r=a; r.rows_per_cta=32;
    float best=1e9;
    for(int rep=0;rep<4;rep++){
      cudaMemcpy(R1,B,elems*4,cudaMemcpyDeviceToDevice); cudaMemcpy(R2,A,elems*4,cudaMemcpyDeviceToDevice);
      cudaEventRecord(e0);
      r.A=A+o0; r.C=R1+o0; r.V=V+o0; r.amp1=a.amp1; k1<4><<<grid,block>>>(r);
      r.A=R1+o0; r.C=R2+o0; r.amp1=a.amp2; k1<4><<<grid,block>>>(r);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms,e0,e1); if(rep>0&&ms<best) best=ms;
    }
    printf("K1x2 (two single-level launches)        %.3f ms / 2 levels  %.1f Gpts/s\n",best,2.0*n*(double)n/(best*1e-3)/1e9);
  }
  // ---- FAST reference: two single-level FAST launches -> RF1, RF2
  float *RF1,*RF2; cudaMalloc(&RF1,elems*4); cudaMalloc(&RF2,elems*4);
  {
    dim3 grid((a.ncol4+255)/256,(n+31)/32), block(256); Args r=a; r.rows_per_cta=32;
    float best=1e9;
    for(int rep=0;rep<4;rep++){
      cudaMemcpy(RF1,B,elems*4,cudaMemcpyDeviceToDevice); cudaMemcpy(RF2,A,elems*4,cudaMemcpyDeviceToDevice);
      cudaEventRecord(e0);
      r.A=A+o0; r.C=RF1+o0; r.V=V+o0; r.amp1=a.amp1; k1f<4><<<grid,block>>>(r);
      r.A=RF1+o0; r.C=RF2+o0; r.amp1=a.amp2; k1f<4><<<grid,block>>>(r);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms,e0,e1); if(rep>0&&ms<best) best=ms;
    }
    printf("K1Fx2 (two single-level FAST launches)  %.3f ms / 2 levels  %.1f Gpts/s\n",best,2.0*n*(double)n/(best*1e-3)/1e9);
  }
  Var vars[]={VWF(8,2,4),VWF(8,2,2),VWF(4,4,4),VWF(8,3,4),VWF(6,4,4),V2F(256,2,1,0),V2F(256,2,0,0),V2F(256,3,0,0),V2F(256,3,1,0),V2F(128,4,1,0),V2F(128,6,0,0),VW(8,2,4,1),VW(8,2,3,1),VW(8,2,2,1),VW(4,4,4,1),VW(2,8,4,1),VW(8,2,4,0),VU(8,2,4,0),VU(8,2,4,1),VU(7,2,4,1),VU(6,2,4,1),VU(5,3,4,1),VU(8,2,3,1),VT(8,2,4),VT(8,2,2),VT(8,2,3),VT(4,4,4),VT(4,4,2),VT(4,3,3),VT(6,2,3),V2(256,2,1,0),V2(256,2,2,2),V2(256,2,2,4),V2(256,2,2,8),V2(256,2,4,2),V2(256,2,4,4),V2(256,2,4,8),V2(256,2,3,4),V2(256,2,5,4),V2(256,2,5,8)};
  const char* filt=argc>2?argv[2]:""; const int rpc_only=argc>3?atoi(argv[3]):0;
  for(auto& vr:vars){
    if(!strstr(vr.name,filt)) continue;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa,vr.f);
    const int nt=vr.nt;
    if(vr.smem) cudaFuncSetAttribute(vr.f,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)vr.smem);
    for(int rpc: {128,256,254,515}){
      if(rpc_only&&rpc!=rpc_only) continue;
      int occ=0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,vr.f,nt,vr.smem);
      int cols_per_blk=(nt/32)*30; dim3 grid((a.ncol4+cols_per_blk-1)/cols_per_blk,(n+rpc-1)/rpc), block(nt);
      Args r=a; r.rows_per_cta=rpc; r.A=A+o0; r.B=B+o0; r.V=V+o0; r.C=C+o0; r.D=D+o0;
      float best=1e9;
      for(int rep=0;rep<4;rep++){
        cudaMemset(C,0,elems*4); cudaMemset(D,0,elems*4);
        cudaEventRecord(e0); vr.f<<<grid,block,vr.smem>>>(r); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms,e0,e1); if(rep>0&&ms<best) best=ms;
      }
      cudaError_t err=cudaGetLastError(); if(err!=cudaSuccess){printf("ERR %s\n",cudaGetErrorString(err));return 1;}
      int s1=-1,s2=-1;
      if(rpc==256||rpc==515){ const bool fast=strstr(vr.name,"K2F")!=nullptr||strstr(vr.name,"K2WF")!=nullptr; s1=cmp(C,fast?RF1:R1,elems,x,y); s2=cmp(D,fast?RF2:R2,elems,x,y); }
      double gp=2.0*n*(double)n/(best*1e-3)/1e9;
      printf("%-16s regs=%3d spill=%zu nt=%3d rpc=%3d occ=%d  %.3f ms / 2 levels  %.1f Gpts/s  (alg %.0f GB/s at 16 B/pt, min-traffic %.0f GB/s at 10 B/pt) same=%d,%d\n",
             vr.name,fa.numRegs,(size_t)fa.localSizeBytes,nt,rpc,occ,best,gp,gp*16,gp*10,s1,s2);
      fflush(stdout);
    }
  }
  return 0;
}
