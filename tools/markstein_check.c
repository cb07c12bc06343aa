#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
static inline uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static uint64_t s=88172645463325252ull;
static inline uint64_t rnd(){s^=s<<13;s^=s>>7;s^=s<<17;return s;}
int main(){
  long bad=0,n=0;
  for(long it=0;it<400000000L;++it){
    uint64_t r=rnd();
    // L in [2^-40,2^40], a in +-[2^-60,2^60]; random mantissas, with adversarial mantissa patterns sometimes
    uint32_t mL=(uint32_t)(r&0x7fffff), ma=(uint32_t)((r>>23)&0x7fffff);
    int sel=(r>>46)&15;
    if(sel==0) mL=0x7fffff; if(sel==1) mL=0; if(sel==2) mL=0x7ffffe; if(sel==3) ma=0x7fffff; if(sel==4) ma=0; if(sel==5) mL=0x400000;
    int eL=127-40+(int)((r>>50)%81), ea=127-60+(int)((r>>57)%121);
    float L=u2f(((uint32_t)eL<<23)|mL), a=u2f(((uint32_t)ea<<23)|ma|((uint32_t)(r>>63)<<31));
    float rl=1.0f/L;                 // correctly rounded reciprocal (IEEE division)
    float q0=a*rl;
    float e=fmaf(-q0,L,a);
    float q=fmaf(e,rl,q0);
    float t=a/L;
    n++;
    if(f2u(q)!=f2u(t)){ if(bad<10) printf("MISMATCH a=%a L=%a q=%a t=%a\n",a,L,q,t); bad++; }
  }
  printf("n=%ld bad=%ld\n",n,bad);
  return 0;
}
