// emulates the lane utilisation of K_roots' Newton loop on the cfg5 workload: items in queue order (pair-major, the two
// brackets of a pair adjacent), warps of 32 items; compares "run to the slowest lane" with "cap + re-queue stragglers"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <vector>
#include <algorithm>
static long g_evals = 0; static double worst_res=0; static long n_bad=0;
#define NRMC_COUNT_EVALS
#include "../nuradiomc_b200/csrc/nrmc_math.cuh"
#include <random>
using namespace nrmc;
int main(){
  IceParams ice; ice.n_ice=1.78; ice.dn=0.423; ice.z0=77; ice.inv_z0=1/77.; ice.inv_dn=1/ice.dn; ice.ns=ice.n_ice-ice.dn; ice.n_refl=0; ice.zr=-1e30; ice.gr=0; ice.nr=ice.n_ice; ice.att_model=0;
  std::mt19937_64 rng(5); std::uniform_real_distribution<double> U(0,1);
  std::vector<int> ev; std::vector<int> piece;
  // vertex-major with 100 antennas per vertex, as the bench
  for(int v=0; v<4000; v++){
    double r=sqrt(U(rng))*6000, ph=U(rng)*2*M_PI, z=-2700*U(rng);
    for(int a=0;a<100;a++){ int st=a/4; double ax=((st%5)-2)*1500., ay=((st/5)-2)*1500., az=-145-5*(a%4);
      Frame2D f; make_frame(r*cos(ph), r*sin(ph), z, ax, ay, az, f);
      PairGeom g; make_pair_geom(ice, f.z1, f.z2, fmax(f.rho,1e-12), g);
      Curve cv; cv.ice=&ice; cv.g=&g; cv.k=0; cv.rcase=1;
      double J1,J2,J3; Bracket br[2]; bool nh; int nb=classify_mode(cv,J1,J2,J3,br,nh);
      if(nh) nb=hump_search(cv,J1,J2,J3,br);
      for(int b=0;b<nb;b++){ g_evals=0; Root rr=solve_bracket(cv,br[b]); ev.push_back((int)g_evals); piece.push_back(br[b].piece);
        double res=fabs(curve_g(cv,br[b].piece,rr.v)); if(res>worst_res) worst_res=res; if(res>1e-8) n_bad++; }
    }
  }
  printf("worst residual |R-rho| %.3e m, residual > 1e-8 m: %ld\n", worst_res, n_bad);
  size_t n=ev.size(); double sum=0; for(int e:ev) sum+=e;
  double cur=0; for(size_t w=0; w<n; w+=32){ int m=0; for(size_t i=w;i<std::min(n,w+32);i++) m=std::max(m,ev[i]); cur+=m*32; }
  printf("roots %zu mean evals %.2f ; current lane-evals per root %.2f (utilisation %.0f%%)\n", n, sum/n, cur/n, 100*sum/cur);
  for(int cap=2; cap<=5; cap++){
    double c=0; std::vector<int> rest;
    for(size_t w=0; w<n; w+=32){ int m=0; for(size_t i=w;i<std::min(n,w+32);i++){ m=std::max(m,std::min(ev[i],cap)); if(ev[i]>cap) rest.push_back(ev[i]-cap);} c+=m*32; }
    for(size_t w=0; w<rest.size(); w+=32){ int m=0; for(size_t i=w;i<std::min(rest.size(),w+32);i++) m=std::max(m,rest[i]); c+=m*32; }
    printf("cap %d: lane-evals per root %.2f (%.0f%% of current), stragglers %.1f%%\n", cap, c/n, 100*c/cur, 100.0*rest.size()/n);
  }
  // thread per ENTRY: bracket 0 of 32 consecutive entries together, then bracket 1 (entries have 2 brackets here)
  { double c=0; size_t ne=n/2; for(size_t w=0; w<ne; w+=32){ for(int j=0;j<2;j++){ int m=0; for(size_t i=w;i<std::min(ne,w+32);i++) m=std::max(m,ev[2*i+j]); c+=m*32; } }
    printf("thread per entry (bracket 0 then 1): lane-evals per root %.2f (%.0f%% of current)\n", c/n, 100*c/cur);
    int h[2][12]={{0}}; for(size_t i=0;i<n;i++) h[i&1][std::min(ev[i],11)]++;
    for(int j=0;j<2;j++){ printf("bracket %d evals histogram:", j); for(int k=0;k<12;k++) printf(" %d:%.1f%%", k, 200.0*h[j][k]/n); printf("\n"); }
    int hp[4][12]={{0}}; long np_[4]={0}; for(size_t i=0;i<n;i++){ hp[piece[i]][std::min(ev[i],11)]++; np_[piece[i]]++; }
    for(int p=0;p<4;p++){ printf("piece %d (%ld):", p, np_[p]); for(int k=0;k<12;k++) if(hp[p][k]) printf(" %d:%.1f%%", k, 100.0*hp[p][k]/np_[p]); printf("\n"); } }
  // sorted by piece
  { std::vector<int> idx(n); for(size_t i=0;i<n;i++) idx[i]=i; std::stable_sort(idx.begin(),idx.end(),[&](int a,int b){return piece[a]<piece[b];});
    double c=0; for(size_t w=0; w<n; w+=32){ int m=0; for(size_t i=w;i<std::min(n,w+32);i++) m=std::max(m,ev[idx[i]]); c+=m*32; }
    printf("sorted by piece: lane-evals per root %.2f (%.0f%% of current)\n", c/n, 100*c/cur); }
}
