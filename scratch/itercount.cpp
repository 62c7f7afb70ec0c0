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
  long cnt[4]={0}, ev[4]={0}; long hist[4][40]={{0}};
  int N=200000;
  for(int i=0;i<N;i++){
    double r=sqrt(U(rng))*6000, ph=U(rng)*2*M_PI, z=-2700*U(rng);
    int st=(int)(U(rng)*25); double ax=((st%5)-2)*1500., ay=((st/5)-2)*1500., az=-145-5*(int)(U(rng)*4);
    Frame2D f; make_frame(r*cos(ph), r*sin(ph), z, ax, ay, az, f);
    PairGeom g; make_pair_geom(ice, f.z1, f.z2, fmax(f.rho,1e-12), g);
    Curve cv; cv.ice=&ice; cv.g=&g; cv.m_dir=mode_coeffs(0,1,false); cv.m_trn=mode_coeffs(0,1,true);
    double J[5]; J[0]=J[4]=-g.rho; J[1]=curve_g(cv,0,1.0); J[3]=curve_g(cv,3,1.0); J[2]=curve_g(cv,1,0.0);
    double pa[4]={0,g.s2max,0,1}, pb[4]={1,0,g.s2max,0};
    for(int p=0;p<4;p++) if((J[p]>0)!=(J[p+1]>0)){ g_evals=0; solve_piece(cv,p,pa[p],J[p],pb[p],J[p+1]); cnt[p]++; ev[p]+=g_evals; hist[p][g_evals>39?39:g_evals]++; }
  }
  for(int p=0;p<4;p++){ printf("piece %d: %ld roots avg %.2f evals | ", p, cnt[p], cnt[p]?(double)ev[p]/cnt[p]:0.); for(int k=0;k<40;k++) if(hist[p][k]) printf("%d:%ld ",k,hist[p][k]); printf("\n"); }
}
