"""CPU: the float-only acosf restatement used on the device (rt_math.cuh: libm_acosf, glibc 2.39
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
