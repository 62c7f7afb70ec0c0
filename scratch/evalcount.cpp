#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
static long g_evals = 0;
#define NRMC_COUNT_EVALS
#include "../nuradiomc_b200/csrc/nrmc_math.cuh"
#include <random>
using namespace nrmc;
int main(){
  IceParams ice; ice.n_ice=1.78; ice.dn=0.423; ice.z0=77; ice.inv_z0=1/77.; ice.ns=ice.n_ice-ice.dn; ice.n_refl=0; ice.zr=-1e30; ice.gr=0; ice.nr=ice.n_ice; ice.att_model=0;
  std::mt19937_64 rng(5); std::uniform_real_distribution<double> U(0,1);
  TraceOutputs o = {0};
  long hist[64]={0}; long nsolh[3]={0}; long ev_by_nsol[3]={0};
  int N=200000;
  for(int i=0;i<N;i++){
    double r=sqrt(U(rng))*6000, ph=U(rng)*2*M_PI, z=-2700*U(rng);
    int st=(int)(U(rng)*25); double ax=((st%5)-2)*1500., ay=((st/5)-2)*1500., az=-145-5*(int)(U(rng)*4);
    g_evals=0;
    int n=trace_pair(ice, r*cos(ph), r*sin(ph), z, ax, ay, az, 0, o, nullptr);
    hist[g_evals>63?63:g_evals]++; nsolh[n]++; ev_by_nsol[n]+=g_evals;
  }
  for(int i=0;i<64;i++) if(hist[i]) printf("%d evals: %ld\n", i, hist[i]);
  for(int n=0;n<3;n++) printf("nsol %d: %ld pairs, avg evals %.2f\n", n, nsolh[n], nsolh[n]? (double)ev_by_nsol[n]/nsolh[n]:0.);
}
