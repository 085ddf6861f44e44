// microbench.cu -- issue-rate probes for the pipes the SPH pair loop uses (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu ; run on the B200 box.
//
// Every probe is a loop of `asm volatile` instructions on independent register chains, so neither nvcc nor ptxas can fold,
// hoist or drop them; the number of target instructions per loop turn is fixed by construction (checked once with
// `cuobjdump -sass`, see tools/README.md) and the loop overhead (one IADD + one branch per UNROLL*NCHAIN target instructions) is
// reported as part of the denominator, not hidden.  Rates are warp-instructions per clock per SM from clock64() spans measured
// INSIDE the kernel by every warp (max end - min start per CTA, averaged over CTAs), with 4 CTAs x 8 warps resident per SM,
// so the figure cannot exceed the 4 issue slots of an SM; the wall-clock time of the launch is printed beside it as a
// cross-check (cycles / ms = the SM clock the run saw).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 8192
#define UNROLL 8
#define CHECK(x) do{ cudaError_t e=(x); if(e!=cudaSuccess){ printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1; } }while(0)

typedef unsigned long long u64;

enum { M_FFMA=0, M_FFMA2, M_FADD2, M_FMUL2, M_FADD, M_HFMA2, M_RCP, M_RSQ, M_FFMA2_ALU, M_FFMA2_MUFU, M_LDS128_BC, M_LDS128_SEQ, M_LDS128_RND,
       M_LDS64_RND, M_LDS32_RND, M_LDS128_NEAR, M_FFMA2_LDS, M_LOP3, M_FSETP, M_COUNT };

template<int MODE> __global__ void __launch_bounds__(256) k_bench(float* out, float seed, u64* span){
  extern __shared__ float4 smem[];
  const int tid=threadIdx.x;
  for(int i=tid;i<2048;i+=blockDim.x)smem[i]=make_float4(seed+i,seed*i,seed-i,1.f+i);
  __shared__ u64 tmin, tmax;
  if(tid==0){ tmin=~0ull; tmax=0ull; }
  __syncthreads();
  float a0=seed+tid,a1=seed*2+tid,a2=seed*3,a3=seed*4,a4=seed*5,a5=seed*6,a6=seed*7,a7=seed*8;
  float b=1.0001f+seed*1e-9f,c=0.0001f*seed;
  u64 p0,p1,p2,p3,p4,p5,p6,p7,pb,pc;
  asm("mov.b64 %0,{%1,%2};":"=l"(p0):"f"(a0),"f"(a1)); asm("mov.b64 %0,{%1,%2};":"=l"(p1):"f"(a2),"f"(a3));
  asm("mov.b64 %0,{%1,%2};":"=l"(p2):"f"(a4),"f"(a5)); asm("mov.b64 %0,{%1,%2};":"=l"(p3):"f"(a6),"f"(a7));
  asm("mov.b64 %0,{%1,%2};":"=l"(p4):"f"(a1),"f"(a0)); asm("mov.b64 %0,{%1,%2};":"=l"(p5):"f"(a3),"f"(a2));
  asm("mov.b64 %0,{%1,%2};":"=l"(p6):"f"(a5),"f"(a4)); asm("mov.b64 %0,{%1,%2};":"=l"(p7):"f"(a7),"f"(a6));
  asm("mov.b64 %0,{%1,%2};":"=l"(pb):"f"(b),"f"(b)); asm("mov.b64 %0,{%1,%2};":"=l"(pc):"f"(c),"f"(c));
  uint32_t u0=tid*2654435761u,u1=u0^0x9e3779b9u,u2=u0+77u,u3=u1+1234567u,u4=u0*3u,u5=u1*5u,u6=u2*7u,u7=u3*9u;
  uint32_t h0=__float_as_uint(a0),h1=__float_as_uint(a1),h2=__float_as_uint(a2),h3=__float_as_uint(a3),h4=h0^h1,h5=h1^h2,h6=h2^h3,h7=h3^h0;
  const uint32_t hb=0x3c003c00u, hc=0x00010001u;
  const uint32_t sbase=(uint32_t)__cvta_generic_to_shared(smem);
  uint32_t adr_bc=sbase, adr_seq=sbase+tid*16u;
  u64 t0=clock64();
  #pragma unroll 1
  for(int i=0;i<ITERS;i++){
    #pragma unroll
    for(int u=0;u<UNROLL;u++){
      if(MODE==M_FFMA){
        asm volatile("fma.rn.f32 %0,%0,%1,%2;":"+f"(a0):"f"(b),"f"(c)); asm volatile("fma.rn.f32 %0,%0,%1,%2;":"+f"(a1):"f"(b),"f"(c));
        asm volatile("fma.rn.f32 %0,%0,%1,%2;":"+f"(a2):"f"(b),"f"(c)); asm volatile("fma.rn.f32 %0,%0,%1,%2;":"+f"(a3):"f"(b),"f"(c));
        asm volatile("fma.rn.f32 %0,%0,%1,%2;":"+f"(a4):"f"(b),"f"(c)); asm volatile("fma.rn.f32 %0,%0,%1,%2;":"+f"(a5):"f"(b),"f"(c));
        asm volatile("fma.rn.f32 %0,%0,%1,%2;":"+f"(a6):"f"(b),"f"(c)); asm volatile("fma.rn.f32 %0,%0,%1,%2;":"+f"(a7):"f"(b),"f"(c));
      }
      if(MODE==M_FADD){
        asm volatile("add.rn.f32 %0,%0,%1;":"+f"(a0):"f"(c)); asm volatile("add.rn.f32 %0,%0,%1;":"+f"(a1):"f"(c));
        asm volatile("add.rn.f32 %0,%0,%1;":"+f"(a2):"f"(c)); asm volatile("add.rn.f32 %0,%0,%1;":"+f"(a3):"f"(c));
        asm volatile("add.rn.f32 %0,%0,%1;":"+f"(a4):"f"(c)); asm volatile("add.rn.f32 %0,%0,%1;":"+f"(a5):"f"(c));
        asm volatile("add.rn.f32 %0,%0,%1;":"+f"(a6):"f"(c)); asm volatile("add.rn.f32 %0,%0,%1;":"+f"(a7):"f"(c));
      }
      if(MODE==M_FFMA2 || MODE==M_FFMA2_ALU || MODE==M_FFMA2_MUFU || MODE==M_FFMA2_LDS){
        asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p0):"l"(pb),"l"(pc)); asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p1):"l"(pb),"l"(pc));
        asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p2):"l"(pb),"l"(pc)); asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p3):"l"(pb),"l"(pc));
        if(MODE==M_FFMA2_ALU){   // 4 FFMA2 + 4 LOP3: do the FMA and ALU pipes overlap inside the 1 instruction/clk/SMSP issue limit?
          asm volatile("lop3.b32 %0,%0,%1,%2,0x96;":"+r"(u0):"r"(u1),"r"(u2)); asm volatile("lop3.b32 %0,%0,%1,%2,0x96;":"+r"(u3):"r"(u1),"r"(u2));
          asm volatile("lop3.b32 %0,%0,%1,%2,0x96;":"+r"(u4):"r"(u1),"r"(u2)); asm volatile("lop3.b32 %0,%0,%1,%2,0x96;":"+r"(u5):"r"(u1),"r"(u2));
        }
        else if(MODE==M_FFMA2_MUFU){   // 4 FFMA2 + 1 MUFU (the ratio of the pair body)
          asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a0));
        }
        else if(MODE==M_FFMA2_LDS){    // 4 FFMA2 + 1 random LDS.128
          float4 v; u0=u0*1664525u+1013904223u;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3},[%4];":"=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w):"r"(sbase+((u0>>8)&2047u)*16u));
          a4+=v.x; a5+=v.y; a6+=v.z; a7+=v.w;
        }
        else{
          asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p4):"l"(pb),"l"(pc)); asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p5):"l"(pb),"l"(pc));
          asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p6):"l"(pb),"l"(pc)); asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p7):"l"(pb),"l"(pc));
        }
      }
      if(MODE==M_FADD2){
        asm volatile("add.rn.f32x2 %0,%0,%1;":"+l"(p0):"l"(pc)); asm volatile("add.rn.f32x2 %0,%0,%1;":"+l"(p1):"l"(pc));
        asm volatile("add.rn.f32x2 %0,%0,%1;":"+l"(p2):"l"(pc)); asm volatile("add.rn.f32x2 %0,%0,%1;":"+l"(p3):"l"(pc));
        asm volatile("add.rn.f32x2 %0,%0,%1;":"+l"(p4):"l"(pc)); asm volatile("add.rn.f32x2 %0,%0,%1;":"+l"(p5):"l"(pc));
        asm volatile("add.rn.f32x2 %0,%0,%1;":"+l"(p6):"l"(pc)); asm volatile("add.rn.f32x2 %0,%0,%1;":"+l"(p7):"l"(pc));
      }
      if(MODE==M_FMUL2){
        asm volatile("mul.rn.f32x2 %0,%0,%1;":"+l"(p0):"l"(pb)); asm volatile("mul.rn.f32x2 %0,%0,%1;":"+l"(p1):"l"(pb));
        asm volatile("mul.rn.f32x2 %0,%0,%1;":"+l"(p2):"l"(pb)); asm volatile("mul.rn.f32x2 %0,%0,%1;":"+l"(p3):"l"(pb));
        asm volatile("mul.rn.f32x2 %0,%0,%1;":"+l"(p4):"l"(pb)); asm volatile("mul.rn.f32x2 %0,%0,%1;":"+l"(p5):"l"(pb));
        asm volatile("mul.rn.f32x2 %0,%0,%1;":"+l"(p6):"l"(pb)); asm volatile("mul.rn.f32x2 %0,%0,%1;":"+l"(p7):"l"(pb));
      }
      if(MODE==M_HFMA2){
        asm volatile("fma.rn.f16x2 %0,%0,%1,%2;":"+r"(h0):"r"(hb),"r"(hc)); asm volatile("fma.rn.f16x2 %0,%0,%1,%2;":"+r"(h1):"r"(hb),"r"(hc));
        asm volatile("fma.rn.f16x2 %0,%0,%1,%2;":"+r"(h2):"r"(hb),"r"(hc)); asm volatile("fma.rn.f16x2 %0,%0,%1,%2;":"+r"(h3):"r"(hb),"r"(hc));
        asm volatile("fma.rn.f16x2 %0,%0,%1,%2;":"+r"(h4):"r"(hb),"r"(hc)); asm volatile("fma.rn.f16x2 %0,%0,%1,%2;":"+r"(h5):"r"(hb),"r"(hc));
        asm volatile("fma.rn.f16x2 %0,%0,%1,%2;":"+r"(h6):"r"(hb),"r"(hc)); asm volatile("fma.rn.f16x2 %0,%0,%1,%2;":"+r"(h7):"r"(hb),"r"(hc));
      }
      if(MODE==M_RCP){   // ptxas folds rcp(rcp(x)): every chain alternates rcp and ex2
        if(u&1){
          asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a0)); asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a1));
          asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a2)); asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a3));
          asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a4)); asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a5));
          asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a6)); asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a7));
        }
        else{
          asm volatile("ex2.approx.ftz.f32 %0,%0;":"+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0,%0;":"+f"(a1));
          asm volatile("ex2.approx.ftz.f32 %0,%0;":"+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0,%0;":"+f"(a3));
          asm volatile("ex2.approx.ftz.f32 %0,%0;":"+f"(a4)); asm volatile("ex2.approx.ftz.f32 %0,%0;":"+f"(a5));
          asm volatile("ex2.approx.ftz.f32 %0,%0;":"+f"(a6)); asm volatile("ex2.approx.ftz.f32 %0,%0;":"+f"(a7));
        }
      }
      if(MODE==M_RSQ){
        asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a0)); asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a1));
        asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a2)); asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a3));
        asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a4)); asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a5));
        asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a6)); asm volatile("rsqrt.approx.ftz.f32 %0,%0;":"+f"(a7));
      }
      if(MODE==M_LOP3){
        asm volatile("lop3.b32 %0,%0,%1,%2,0x96;":"+r"(u0):"r"(u6),"r"(u7)); asm volatile("lop3.b32 %0,%0,%1,%2,0x96;":"+r"(u1):"r"(u6),"r"(u7));
        asm volatile("lop3.b32 %0,%0,%1,%2,0x96;":"+r"(u2):"r"(u6),"r"(u7)); asm volatile("lop3.b32 %0,%0,%1,%2,0x96;":"+r"(u3):"r"(u6),"r"(u7));
        asm volatile("lop3.b32 %0,%0,%1,%2,0x96;":"+r"(u4):"r"(u6),"r"(u7)); asm volatile("lop3.b32 %0,%0,%1,%2,0x96;":"+r"(u5):"r"(u6),"r"(u7));
        asm volatile("lop3.b32 %0,%0,%1,%2,0x96;":"+r"(h0):"r"(u6),"r"(u7)); asm volatile("lop3.b32 %0,%0,%1,%2,0x96;":"+r"(h1):"r"(u6),"r"(u7));
      }
      if(MODE==M_FSETP){   // compare + predicated integer add: 8 x (FSETP + @p IADD)
        asm volatile("{.reg .pred p; setp.lt.f32 p,%1,%2; @p add.u32 %0,%0,1;}":"+r"(u0):"f"(a0),"f"(c)); asm volatile("{.reg .pred p; setp.lt.f32 p,%1,%2; @p add.u32 %0,%0,1;}":"+r"(u1):"f"(a1),"f"(c));
        asm volatile("{.reg .pred p; setp.lt.f32 p,%1,%2; @p add.u32 %0,%0,1;}":"+r"(u2):"f"(a2),"f"(c)); asm volatile("{.reg .pred p; setp.lt.f32 p,%1,%2; @p add.u32 %0,%0,1;}":"+r"(u3):"f"(a3),"f"(c));
        asm volatile("{.reg .pred p; setp.lt.f32 p,%1,%2; @p add.u32 %0,%0,1;}":"+r"(u4):"f"(a4),"f"(c)); asm volatile("{.reg .pred p; setp.lt.f32 p,%1,%2; @p add.u32 %0,%0,1;}":"+r"(u5):"f"(a5),"f"(c));
        asm volatile("{.reg .pred p; setp.lt.f32 p,%1,%2; @p add.u32 %0,%0,1;}":"+r"(u6):"f"(a6),"f"(c)); asm volatile("{.reg .pred p; setp.lt.f32 p,%1,%2; @p add.u32 %0,%0,1;}":"+r"(u7):"f"(a7),"f"(c));
      }
      if(MODE==M_LDS128_BC || MODE==M_LDS128_SEQ){   // 8 loads per turn, addresses advance by a register stride (no address math inside)
        const uint32_t ad=(MODE==M_LDS128_BC? adr_bc: adr_seq);
        float4 v;
        #define LDSK(K) asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3},[%4+" #K "];":"=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w):"r"(ad)); a0+=v.x; a1+=v.y; a2+=v.z; a3+=v.w;
        LDSK(0) LDSK(512) LDSK(1024) LDSK(1536) LDSK(2048) LDSK(2560) LDSK(3072) LDSK(3584)
        #undef LDSK
      }
      if(MODE==M_LDS128_RND || MODE==M_LDS64_RND || MODE==M_LDS32_RND || MODE==M_LDS128_NEAR){   // 4 loads per turn at per-lane pseudo-random 16-byte slots (+ the LCG: IMAD, SHF, LOP, LEA)
        #pragma unroll
        for(int k=0;k<4;k++){
          u0=u0*1664525u+1013904223u;
          // NEAR: the lanes of a warp read inside one window of 64 slots (1 KB) that moves with the turn: the pattern of hit lists of neighbouring targets
          const uint32_t slot=(MODE==M_LDS128_NEAR? ((uint32_t(i)*37u+(u0>>26))&2047u): ((u0>>8)&2047u));
          const uint32_t ad=sbase+slot*16u;
          float4 v=make_float4(0,0,0,0);
          if(MODE==M_LDS128_RND || MODE==M_LDS128_NEAR){ asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3},[%4];":"=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w):"r"(ad)); a0+=v.x; a1+=v.y; a2+=v.z; a3+=v.w; }
          if(MODE==M_LDS64_RND){ asm volatile("ld.shared.v2.f32 {%0,%1},[%2];":"=f"(v.x),"=f"(v.y):"r"(ad)); a0+=v.x; a1+=v.y; }
          if(MODE==M_LDS32_RND){ asm volatile("ld.shared.f32 %0,[%1];":"=f"(v.x):"r"(ad)); a0+=v.x; }
        }
      }
    }
  }
  u64 t1=clock64();
  atomicMin(&tmin,t0); atomicMax(&tmax,t1);
  __syncthreads();
  out[blockIdx.x*blockDim.x+tid]=a0+a1+a2+a3+a4+a5+a6+a7+float(u0^u1^u2^u3^u4^u5^u6^u7)+float(h0^h1^h2^h3^h4^h5^h6^h7)+float(p0^p1^p2^p3^p4^p5^p6^p7);
  if(tid==0)span[blockIdx.x]=tmax-tmin;
}

static double g_smclk_khz=1965000.0;
struct Probe { const char* name; double target_per_turn; double other_per_turn; };

template<int MODE> static int run(const char* name,double target,double other,int nsm,float* dout,u64* dspan){
  const int threads=256, blocks=nsm*4;   // 4 CTAs x 8 warps per SM = 32 warps/SM = 8 per scheduler
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_bench<MODE><<<blocks,threads,2048*16>>>(dout,1.0f,dspan);
  CHECK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  k_bench<MODE><<<blocks,threads,2048*16>>>(dout,1.0f,dspan);
  cudaEventRecord(e1);
  CHECK(cudaDeviceSynchronize());
  float ms=0; cudaEventElapsedTime(&ms,e0,e1);
  static u64 hs[8192]; cudaMemcpy(hs,dspan,sizeof(u64)*blocks,cudaMemcpyDeviceToHost);
  double cyc=0; for(int i=0;i<blocks;i++)cyc+=double(hs[i]); cyc/=blocks;
  // per SM: 4 CTAs x 8 warps, each ITERS turns of UNROLL x (target + other) instructions + 2 of loop control
  const double warps=4.0*(threads/32);
  const double wt=double(ITERS)*UNROLL*target*warps, wall=double(ITERS)*(UNROLL*(target+other)+2.0)*warps;
  // clock64() does not tick at the SM clock on this part (its rate against the event time changes from probe to probe), so the
  // rates are taken from the event time of the whole launch at the SM clock the device reports (launch overhead included: a lower bound)
  const double smclk=g_smclk_khz*1e3, clk=ms*1e-3*smclk;
  printf("%-46s %8.3f ms  target %6.3f  all %6.3f warp-inst/clk/SM   [clock64 span %.0f = %.2f GHz]\n",name,ms,wt/clk,wall/clk,cyc,cyc/(ms*1e6));
  return 0;
}

int main(){
  int dev=0; cudaDeviceProp pr; CHECK(cudaGetDeviceProperties(&pr,dev));
  printf("device: %s  SMs=%d  clock=%d kHz\n",pr.name,pr.multiProcessorCount,pr.clockRate);
  printf("rates: 'target' counts the probed instruction only, 'all' adds the companions and loop control; an SM issues at most 4 warp-inst/clk\n");
  const int nsm=pr.multiProcessorCount; g_smclk_khz=pr.clockRate;
  float* dout; u64* dspan;
  CHECK(cudaMalloc(&dout,sizeof(float)*nsm*4*256)); CHECK(cudaMalloc(&dspan,sizeof(u64)*8192));
  run<M_FFMA>("FFMA (8 chains)",8,0,nsm,dout,dspan);
  run<M_FADD>("FADD (8 chains)",8,0,nsm,dout,dspan);
  run<M_FFMA2>("FFMA2 = fma.rn.f32x2 (8 chains)",8,0,nsm,dout,dspan);
  run<M_FADD2>("FADD2 = add.rn.f32x2 (8 chains)",8,0,nsm,dout,dspan);
  run<M_FMUL2>("FMUL2 = mul.rn.f32x2 (8 chains)",8,0,nsm,dout,dspan);
  run<M_HFMA2>("HFMA2 = fma.rn.f16x2 (8 chains)",8,0,nsm,dout,dspan);
  run<M_LOP3>("LOP3 (8 chains)",8,0,nsm,dout,dspan);
  run<M_FSETP>("FSETP + predicated IADD (8 pairs)",8,8,nsm,dout,dspan);
  run<M_RSQ>("MUFU.RSQ (8 chains)",8,0,nsm,dout,dspan);
  run<M_FFMA2_ALU>("4 FFMA2 + 4 LOP3",4,4,nsm,dout,dspan);
  run<M_FFMA2_MUFU>("4 FFMA2 + 1 MUFU.RSQ",4,1,nsm,dout,dspan);
  run<M_FFMA2_LDS>("4 FFMA2 + 1 LDS.128 random (+4 addr +4 FADD)",4,9,nsm,dout,dspan);
  run<M_LDS128_BC>("LDS.128 broadcast (8 + 32 FADD)",8,32,nsm,dout,dspan);
  run<M_LDS128_SEQ>("LDS.128 lane-consecutive (8 + 32 FADD)",8,32,nsm,dout,dspan);
  run<M_LDS128_RND>("LDS.128 random 16B slots in 32 KB (4 + 4x8)",4,32,nsm,dout,dspan);
  run<M_LDS128_NEAR>("LDS.128 random inside a 1 KB window (4 + 4x10)",4,40,nsm,dout,dspan);
  run<M_LDS64_RND>("LDS.64 random 16B slots (4 + 4x6)",4,24,nsm,dout,dspan);
  run<M_LDS32_RND>("LDS.32 random 16B slots (4 + 4x5)",4,20,nsm,dout,dspan);
  return 0;
}
