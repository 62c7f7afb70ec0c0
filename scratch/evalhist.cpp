// evaluations per solve_piece call by piece and by bracket index (cfg5 geometry): is the iteration count systematic?
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
static long g_evals = 0;
#define NRMC_COUNT_EVALS
#include "../nuradiomc_b200/csrc/nrmc_math.cuh"
#include <random>
using namespace nrmc;
int main(int argc, char **argv){
  double dn = argc>1?atof(argv[1]):0.423, z0 = argc>2?atof(argv[2]):77., amin = argc>3?atof(argv[3]):-145., aspan = argc>4?atof(argv[4]):-15., rmax = argc>5?atof(argv[5]):6000., zmin = argc>6?atof(argv[6]):-2700.;
  IceParams ice; ice.n_ice=1.78; ice.dn=dn; ice.z0=z0; ice.inv_z0=1/z0; ice.inv_dn=1/dn; ice.ns=ice.n_ice-ice.dn; ice.n_refl=0; ice.zr=-1e30; ice.gr=0; ice.nr=ice.n_ice; ice.att_model=0;
  std::mt19937_64 rng(5); std::uniform_real_distribution<double> U(0,1);
  long hist[2][4][16]={{{0}}}; long cnt[2][4]={{0}}, sum[2][4]={{0}};
  int N=300000;
  for(int i=0;i<N;i++){
    double r=sqrt(U(rng))*rmax, ph=U(rng)*2*M_PI, z=zmin*U(rng);
    int st=(int)(U(rng)*25); double ax=((st%5)-2)*1500., ay=((st/5)-2)*1500., az=amin+aspan*U(rng);
    Frame2D f; make_frame(r*cos(ph), r*sin(ph), z, ax, ay, az, f);
    PairGeom g; make_pair_geom(ice, f.z1, f.z2, fmax(f.rho,1e-12), g);
    Curve cv; cv.ice=&ice; cv.g=&g; cv.k=0; cv.rcase=1;
    double J1,J2,J3; Bracket br[2]; bool need_hump;
    int nb = classify_mode(cv, J1, J2, J3, br, need_hump);
    if (need_hump) nb = hump_search(cv, J1, J2, J3, br);
    for (int b=0;b<nb;b++){ g_evals=0; Root rt=solve_bracket(cv, br[b]); int e=g_evals>15?15:g_evals; hist[b][br[b].piece][e]++; cnt[b][br[b].piece]++; sum[b][br[b].piece]+=g_evals; }
  }
  for(int b=0;b<2;b++) for(int p=0;p<4;p++) if(cnt[b][p]){ printf("bracket %d piece %d: %8ld solves, mean %.2f evals | ", b, p, cnt[b][p], (double)sum[b][p]/cnt[b][p]); for(int e=1;e<12;e++) printf("%d:%.1f%% ", e, 100.0*hist[b][p][e]/cnt[b][p]); printf("\n"); }
}
