"""CPU: libm restatements used on the device against the host libm the reference links.
sinf / cosf / atanf / atan2f (environment maps) are extracted from rt_math.cuh and compared on tens of millions of
arguments of the range the path produces; acosf: the float-only acosf restatement used on the device (rt_math.cuh: libm_acosf, glibc 2.39
sysdeps/ieee754/flt-32/e_acosf.c) against the host libm the reference links, on a dense sample of [-1, 1] plus edge
cases.  (The full 2^31-point sweep was run once when the port was written: 0 mismatches.)"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_device_acosf_restatement_matches_host_libm(tmp_path):
    src = open(os.path.join(ROOT, "ray_b200", "csrc", "rt_math.cuh")).read()
    m = re.search(r"RT_FN float libm_acosf\(float x\) \{(.*?)\n\}\n", src, re.S)
    assert m, "libm_acosf not found in rt_math.cuh"
    body = m.group(1)
    c = r'''
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static inline int __float_as_int(float f){int i;memcpy(&i,&f,4);return i;}
static inline float __int_as_float(int i){float f;memcpy(&f,&i,4);return f;}
static float libm_acosf(float x) {''' + body + r'''
}
int main(void){
  long bad=0,n=0;
  for (uint32_t u=0; u<=0x3f800000u; u+=97) { for(int s=0;s<2;s++){
      float x=__int_as_float((int)(u|(s?0x80000000u:0))); volatile float xv=x;
      float a=acosf(xv), b=libm_acosf(x); n++;
      if (__float_as_int(a)!=__float_as_int(b)) { if(bad<5) printf("x=%a libm=%a port=%a\n",x,a,b); bad++; } } }
  const float edge[] = {1.0f,-1.0f,0.0f,-0.0f,0.5f,-0.5f,0.49999997f,0.50000006f,1e-30f,0x1p-60f,0.99999994f,-0.99999994f};
  for (unsigned i=0;i<sizeof(edge)/sizeof(edge[0]);i++){ volatile float xv=edge[i]; float a=acosf(xv), b=libm_acosf(edge[i]); n++;
      if (__float_as_int(a)!=__float_as_int(b)) { printf("edge x=%a libm=%a port=%a\n",edge[i],a,b); bad++; } }
  printf("n=%ld bad=%ld\n",n,bad); return bad!=0; }
'''
    f = tmp_path / "acos.c"
    f.write_text(c)
    exe = tmp_path / "acos"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(f), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout


def test_device_trig_restatements_match_host_libm(tmp_path):
    """libm_sinf / libm_cosf / libm_atanf / libm_atan2f (rt_math.cuh), compiled as C, against sinf / cosf / atanf / atan2f."""
    src = open(os.path.join(ROOT, "ray_b200", "csrc", "rt_math.cuh")).read()
    i0 = src.index("RT_FN float libm_sincosf_poly(")
    i1 = src.index("// expf of the host libm (glibc 2.39")
    code = src[i0:i1].replace("RT_FN", "static")
    c = r'''
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdbool.h>
static inline int __float_as_int(float f){int i;memcpy(&i,&f,4);return i;}
static inline uint32_t __float_as_uint(float f){uint32_t i;memcpy(&i,&f,4);return i;}
static inline float __int_as_float(int i){float f;memcpy(&f,&i,4);return f;}
static inline int __double2int_rz(double d){return (int)d;}
#define double(x) ((double)(x))
#define float(x) ((float)(x))
''' + code + r'''
static uint32_t rng=12345; static uint32_t r32(void){ rng^=rng<<13; rng^=rng>>17; rng^=rng<<5; return rng; }
static double u01(void){ return (double)r32()/4294967296.0; }
int main(void){ long bad=0;
  for(long i=0;i<8000000;i++){
    float x=(float)(u01()*16.0-8.0); volatile float xv=x;
    if(__float_as_int(sinf(xv))!=__float_as_int(libm_sinf(x))){ if(bad<5)printf("sinf %a\n",x); bad++; }
    if(__float_as_int(cosf(xv))!=__float_as_int(libm_cosf(x))){ if(bad<5)printf("cosf %a\n",x); bad++; }
    float t=(float)(u01()*60.0-30.0); volatile float tv=t;
    if(__float_as_int(atanf(tv))!=__float_as_int(libm_atanf(t))){ if(bad<5)printf("atanf %a\n",t); bad++; }
    float yy=(float)(u01()*2.0-1.0), xx=(float)(u01()*2.0-1.0); volatile float yv=yy, xw=xx;
    if(__float_as_int(atan2f(yv,xw))!=__float_as_int(libm_atan2f(yy,xx))){ if(bad<5)printf("atan2f %a %a\n",yy,xx); bad++; } }
  const float e[]={0.0f,-0.0f,1.0f,-1.0f,0x1p-13f,0x1p-12f,0.78539813f,0.7853982f,0.78539824f,3.1415927f,6.2831855f,6.9831855f,-6.9831855f,1e-30f,0.4375f,0.6875f,1.1875f,2.4375f,1e10f};
  for(unsigned i=0;i<sizeof(e)/sizeof(e[0]);i++) for(unsigned j=0;j<sizeof(e)/sizeof(e[0]);j++){ volatile float a=e[i], b=e[j];
    if(__float_as_int(atan2f(a,b))!=__float_as_int(libm_atan2f(e[i],e[j]))){ printf("atan2f edge %a %a\n",e[i],e[j]); bad++; }
    if(i==0 && fabsf(e[j])<100.0f){ if(__float_as_int(sinf(b))!=__float_as_int(libm_sinf(e[j]))){ printf("sinf edge %a\n",e[j]); bad++; }
      if(__float_as_int(cosf(b))!=__float_as_int(libm_cosf(e[j]))){ printf("cosf edge %a\n",e[j]); bad++; } } }
  printf("bad=%ld\n",bad); return bad!=0; }
'''
    f = tmp_path / "trig.c"
    f.write_text(c)
    exe = tmp_path / "trig"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(f), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout


def test_device_expf_restatement_matches_host_libm(tmp_path):
    """libm_expf (rt_math.cuh; the NLM denoiser's weights), compiled as C, against the host expf on the argument range
    the filter produces (<= 0) plus positive and overflow / underflow edge cases."""
    src = open(os.path.join(ROOT, "ray_b200", "csrc", "rt_math.cuh")).read()
    i0 = src.index("RT_FN float libm_expf(float x) {")
    i1 = src.index("// logf and powf of the host libm (glibc 2.39")
    code = src[i0:i1].replace("RT_FN", "static")
    c = r'''
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static inline uint32_t __float_as_uint(float f){uint32_t i;memcpy(&i,&f,4);return i;}
static inline float __uint_as_float(uint32_t i){float f;memcpy(&f,&i,4);return f;}
static inline long long __double_as_longlong(double d){long long i;memcpy(&i,&d,8);return i;}
static inline double __longlong_as_double(long long i){double d;memcpy(&d,&i,8);return d;}
#define double(x) ((double)(x))
#define float(x) ((float)(x))
#define uint64_t(x) ((uint64_t)(x))
''' + code + r'''
static uint32_t rng=777; static uint32_t r32(void){ rng^=rng<<13; rng^=rng>>17; rng^=rng<<5; return rng; }
int main(void){ long bad=0;
  for(long i=0;i<20000000;i++){ float x=-(float)((double)r32()/4294967296.0*112.0); if(i&1) x=-(float)((double)r32()/4294967296.0*4.0); if((i&15)==3) x=-x*20.0f;
    volatile float xv=x; float a=expf(xv), b=libm_expf(x); if(__float_as_uint(a)!=__float_as_uint(b)){ if(bad<5)printf("x=%a %a %a\n",x,a,b); bad++; } }
  const float e[]={0.0f,-0.0f,88.0f,88.72f,89.0f,-87.0f,-88.0f,-100.0f,-103.9f,-104.0f,-150.0f,-1e30f,1e-30f,-1e-30f};
  for(unsigned i=0;i<sizeof(e)/sizeof(e[0]);i++){ volatile float xv=e[i]; if(__float_as_uint(expf(xv))!=__float_as_uint(libm_expf(e[i]))){ printf("edge %a\n",e[i]); bad++; } }
  printf("bad=%ld\n",bad); return bad!=0; }
'''
    f = tmp_path / "expf.c"
    f.write_text(c)
    exe = tmp_path / "expf"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(f), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout


def test_device_logf_powf_restatements_match_host_libm(tmp_path):
    """libm_logf (D_GTR1 of the clearcoat lobe) and libm_powf (sRGB OETF / 1/gamma of the display transform), compiled as
    C, against the host logf / powf on the argument ranges the path produces plus edge cases."""
    src = open(os.path.join(ROOT, "ray_b200", "csrc", "rt_math.cuh")).read()
    i0 = src.index("struct LibmLogTab {")
    i1 = src.index("// exp2f(float(e) - 128.0f) of rgbe_to_rgb")
    code = src[i0:i1].replace("RT_FN", "static")
    c = r'''
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static inline uint32_t __float_as_uint(float f){uint32_t i;memcpy(&i,&f,4);return i;}
static inline float __uint_as_float(uint32_t i){float f;memcpy(&f,&i,4);return f;}
static inline long long __double_as_longlong(double d){long long i;memcpy(&i,&d,8);return i;}
static inline double __longlong_as_double(long long i){double d;memcpy(&d,&i,8);return d;}
#define double(x) ((double)(x))
#define float(x) ((float)(x))
#define int(x) ((int)(x))
#define uint64_t(x) ((uint64_t)(x))
typedef struct LibmLogTab LibmLogTab;
''' + code + r'''
static uint32_t rng=4242; static uint32_t r32(void){ rng^=rng<<13; rng^=rng>>17; rng^=rng<<5; return rng; }
static double u01(void){ return (double)r32()/4294967296.0; }
static int same(float a, float b){ return __float_as_uint(a)==__float_as_uint(b) || (a!=a && b!=b); }
int main(void){ long bad=0;
  const float ys[] = {1.0f/2.4f, 1.0f/2.2f, 2.4f, 0.5f, 1.0f/1.8f, 3.0f, 0.45f};
  for(long i=0;i<12000000;i++){
    float x = (float)(u01()*1.2); if((i&7)==1) x = (float)(u01()*64.0); if((i&7)==2) x = __uint_as_float(r32() & 0x7f7fffffu);
    float y = ys[i % 7]; volatile float xv=x, yv=y;
    if(!same(powf(xv,yv), libm_powf(x,y))){ if(bad<5)printf("powf %a %a: %a %a\n",x,y,powf(xv,yv),libm_powf(x,y)); bad++; }
    float l = (float)(u01()*4.0); if((i&3)==1) l = __uint_as_float(r32() & 0x7f7fffffu);
    volatile float lv=l;
    if(!same(logf(lv), libm_logf(l))){ if(bad<5)printf("logf %a: %a %a\n",l,logf(lv),libm_logf(l)); bad++; } }
  const float e[]={0.0f,1.0f,0x1p-126f,0x1p-127f,0x1p-149f,0.0031308f,1e-30f,1e30f,3.4e38f,0.99999994f,1.0000001f,2.0f,0.5f};
  for(unsigned i=0;i<sizeof(e)/sizeof(e[0]);i++) for(unsigned j=0;j<7;j++){ volatile float a=e[i], b=ys[j];
    if(!same(powf(a,b), libm_powf(e[i],ys[j]))){ printf("powf edge %a %a\n",e[i],ys[j]); bad++; }
    if(j==0 && !same(logf(a), libm_logf(e[i]))){ printf("logf edge %a\n",e[i]); bad++; } }
  printf("bad=%ld\n",bad); return bad!=0; }
'''
    f = tmp_path / "powf.c"
    f.write_text(c)
    exe = tmp_path / "powf"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(f), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout
