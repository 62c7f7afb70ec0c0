// CPU study of the Newton termination thresholds of solve_piece: evaluations per pair and deviation of C0 / path length / travel time
// from the strict setting, on the cfg5 geometry.  g++ -O2 -o /tmp/tolstudy scratch/tolstudy.cpp && /tmp/tolstudy
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
static long g_evals = 0;
static double g_step_tol = 1e-9, g_gtol = 1e-10;
#define NRMC_COUNT_EVALS
#define NRMC_SOLVE_STEP_TOL g_step_tol
#define NRMC_SOLVE_GTOL g_gtol
#include "../nuradiomc_b200/csrc/nrmc_math.cuh"
#include <random>
#include <vector>
using namespace nrmc;
int main(){
  IceParams ice; ice.n_ice=1.78; ice.dn=0.423; ice.z0=77; ice.inv_z0=1/77.; ice.inv_dn=1/0.423; ice.ns=ice.n_ice-ice.dn; ice.n_refl=0; ice.zr=-1e30; ice.gr=0; ice.nr=ice.n_ice; ice.att_model=0;
  const int N=400000;
  std::vector<double> C0ref(2*N), Pref(2*N), Tref(2*N); std::vector<int> nref(N);
  double steps[] = {1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4};
  double gtols[] = {1e-10, 1e-10, 1e-10, 1e-10, 1e-10, 1e-10, 1e-8, 1e-7};
  for (int pass=0; pass<9; ++pass) {
    if (pass < 6) { g_step_tol = steps[pass]; g_gtol = 1e-10; }
    else if (pass == 6) { g_step_tol = 1e-6; g_gtol = 1e-8; }
    else if (pass == 7) { g_step_tol = 1e-6; g_gtol = 1e-7; }
    else { g_step_tol = 1e-5; g_gtol = 1e-7; }
    std::mt19937_64 rng(5); std::uniform_real_distribution<double> U(0,1);
    long ev=0; double dC=0, dP=0, dT=0, dPabs=0; long mism=0;
    for(int i=0;i<N;i++){
      double r=sqrt(U(rng))*6000, ph=U(rng)*2*M_PI, z=-2700*U(rng);
      int st=(int)(U(rng)*25); double ax=((st%5)-2)*1500., ay=((st/5)-2)*1500., az=-145-5*(int)(U(rng)*4);
      int32_t ns=0, status=0; int8_t ty[2], rf[2], rc[2]; double C0[2], C1[2], pl[2], tt[2], la[6], re[6], ra[2];
      TraceOutputs o = {0}; o.n_sol=&ns; o.status=&status; o.type=ty; o.reflection=rf; o.reflection_case=rc; o.C0=C0; o.C1=C1; o.path_length=pl; o.travel_time=tt; o.launch=la; o.receive=re; o.reflection_angle=ra;
      g_evals=0;
      int n=trace_pair(ice, r*cos(ph), r*sin(ph), z, ax, ay, az, 0, o, nullptr);
      ev+=g_evals;
      if (pass==0) { nref[i]=n; for(int s=0;s<n;s++){C0ref[2*i+s]=C0[s];Pref[2*i+s]=pl[s];Tref[2*i+s]=tt[s];} }
      else { if (n!=nref[i]) {mism++; continue;} for(int s=0;s<n;s++){ dC=fmax(dC,fabs(C0[s]/C0ref[2*i+s]-1)); dP=fmax(dP,fabs(pl[s]/Pref[2*i+s]-1)); dPabs=fmax(dPabs,fabs(pl[s]-Pref[2*i+s])); dT=fmax(dT,fabs(tt[s]/Tref[2*i+s]-1)); } }
    }
    printf("step %.0e gtol %.0e: evals/pair %.3f  count mismatches %ld  max rel dC0 %.2e  dPath %.2e (abs %.2e m)  dTime %.2e\n", g_step_tol, g_gtol, (double)ev/N, mism, dC, dP, dPabs, dT);
  }
}
