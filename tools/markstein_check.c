#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
static inline uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static uint64_t s=88172645463325252ull;
static inline uint64_t rnd(){s^=s<<13;s^=s>>7;s^=s<<17;return s;}
// usage: markstein_check [A]   numerators in +-[2^-A, 2^A] (default 84: what csrc/common.cuh safe_factor admits as a product of two
// factors in 2^-42 .. 2^42), divisors in [2^-40, 2^40] (safe_divisor)
int main(int argc,char**argv){
  long bad=0,n=0;
  const int A=argc>1?atoi(argv[1]):84;
  for(long it=0;it<400000000L;++it){
    uint64_t r=rnd();
    // L in [2^-40,2^40], a in +-[2^-A,2^A]; random mantissas, with adversarial mantissa patterns sometimes; one draw in 16 pins an
    // exponent to the end of its range
    uint32_t mL=(uint32_t)(r&0x7fffff), ma=(uint32_t)((r>>23)&0x7fffff);
    int sel=(r>>46)&15;
    if(sel==0) mL=0x7fffff; if(sel==1) mL=0; if(sel==2) mL=0x7ffffe; if(sel==3) ma=0x7fffff; if(sel==4) ma=0; if(sel==5) mL=0x400000;
    const uint64_t r2=rnd();
    int eL=127-40+(int)((r>>50)%81), ea=127-A+(int)(r2%(uint64_t)(2*A+1));
    const int pin=(int)((r2>>32)&63);
    if(pin==0) ea=127-A; if(pin==1) ea=127+A; if(pin==2) eL=127-40; if(pin==3) eL=127+40;
    float L=u2f(((uint32_t)eL<<23)|mL), a=u2f(((uint32_t)ea<<23)|ma|((uint32_t)(r>>63)<<31));
    float rl=1.0f/L;                 // correctly rounded reciprocal (IEEE division)
    float q0=a*rl;
    float e=fmaf(-q0,L,a);
    float q=fmaf(e,rl,q0);
    float t=a/L;
    n++;
    if(f2u(q)!=f2u(t)){ if(bad<10) printf("MISMATCH a=%a L=%a q=%a t=%a\n",a,L,q,t); bad++; }
  }
  printf("A=%d n=%ld bad=%ld\n",A,n,bad);
  return 0;
}
