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
  int shown=0; long hist[4][16]={{0}};
  for(int i=0;i<200000;i++){
    double r=sqrt(U(rng))*6000, ph=U(rng)*2*M_PI, z=-2700*U(rng);
    int st=(int)(U(rng)*25); double ax=((st%5)-2)*1500., ay=((st/5)-2)*1500., az=-145-5*(int)(U(rng)*4);
    Frame2D f; make_frame(r*cos(ph), r*sin(ph), z, ax, ay, az, f);
    PairGeom g; make_pair_geom(ice, f.z1, f.z2, fmax(f.rho,1e-12), g);
    Curve cv; cv.ice=&ice; cv.g=&g; cv.k=0; cv.rcase=1;
    double J1,J2,J3; Bracket br[2]; bool nh; int nb=classify_mode(cv,J1,J2,J3,br,nh);
    for(int b=0;b<nb;b++){ g_evals=0; Root rt=solve_bracket(cv,br[b]); int e=g_evals>15?15:g_evals; hist[br[b].piece][e]++;
      if(br[b].piece==2 && g_evals>=6 && shown<6){ shown++; printf("P2 slow: z1=%.0f z2=%.0f rho=%.1f bracket t[%.4f,%.4f] g[%.3g,%.3g] root t=%.6f evals=%ld\n", g.z1,g.z2,g.rho,br[b].a,br[b].b,br[b].ga,br[b].gb,rt.v,g_evals);
        // sample the function shape
        for(int k=0;k<=10;k++){ double t=br[b].a+(br[b].b-br[b].a)*k/10.0; double d; double gg=curve_gd(cv,2,t,d); printf("   t=%.4f g=%.4g dg=%.4g\n",t,gg,d);} }
    }
  }
  for(int p=0;p<4;p++){ printf("P%d:",p); for(int e=0;e<16;e++) if(hist[p][e]) printf(" %d:%ld",e,hist[p][e]); printf("\n"); }
}
