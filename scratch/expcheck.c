#include <stdio.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
static const double c[10]={1.0/39916800,1.0/3628800,1.0/362880,1.0/40320,1.0/5040,1.0/720,1.0/120,1.0/24,1.0/6,0.5};
static double red(double x,int*k){ double t=fma(x,1.4426950408889634,6755399441055744.0); int64_t b; memcpy(&b,&t,8); *k=(int32_t)(b&0xffffffff); double kd=t-6755399441055744.0; double r=fma(kd,-6.93147180369123816490e-01,x); return fma(kd,-1.90821492927058770002e-10,r);}
static double poly(double r){ double p=c[0]; for(int i=1;i<10;i++) p=fma(p,r,c[i]); return fma(p*r,r,r);}
static double exp_c(double x){ int k; double r=red(fmin(fmax(x,-700),700),&k); double p=poly(r)+1.0; int64_t b; memcpy(&b,&p,8); b+=((int64_t)k<<52); memcpy(&p,&b,8); return p;}
static double expm1_c(double x){ int k; double r=red(fmin(fmax(x,-700),700),&k); double q=poly(r); int64_t b=((int64_t)(1023+k))<<52; double s; memcpy(&s,&b,8); return fma(s,q,s-1.0);}
int main(){ double me=0,mm=0; for(int i=0;i<2000000;i++){ double x=-60.0+ 61.0*i/2000000.0; long double e=expl((long double)x); double r1=fabs((double)((exp_c(x)-e)/e)); if(r1>me)me=r1; if(x<0){ long double m=expm1l((long double)x); double r2=fabs((double)((expm1_c(x)-m)/m)); if(r2>mm)mm=r2; } }
 for(int i=1;i<400;i++){ double x=-pow(10.0,-i/20.0); long double m=expm1l((long double)x); double r2=fabs((double)((expm1_c(x)-m)/m)); if(r2>mm)mm=r2; }
 printf("max rel err exp_c %.3e expm1_c %.3e ; exp_c(-745)=%g exp_c(-1e9)=%g\n",me,mm,exp_c(-745),exp_c(-1e9)); }
