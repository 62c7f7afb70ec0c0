#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include "../nuradiomc_b200/csrc/nrmc_math.cuh"
using namespace nrmc;
int main(){
  IceParams ice; ice.n_ice=1.78; ice.dn=0.51; ice.z0=37.25; ice.inv_z0=1/37.25; ice.ns=ice.n_ice-ice.dn; ice.n_refl=0; ice.zr=-1e30; ice.gr=0; ice.nr=ice.n_ice; ice.att_model=0;
  double X1[3]={-9180.4207706, 1164.20989001, -262.20068663}, X2[3]={0,0,-404.48022707};
  Frame2D f; make_frame(X1[0],X1[1],X1[2],X2[0],X2[1],X2[2],f);
  PairGeom g; make_pair_geom(ice,f.z1,f.z2,f.rho,g);
  printf("z1=%g z2=%g rho=%g n1=%.12f n2=%.12f tmin=%g s2max=%g\n", g.z1,g.z2,g.rho,g.n1,g.n2,g.tmin,g.s2max);
  Curve cv; cv.ice=&ice; cv.g=&g; cv.k=0; cv.rcase=1;
  double J1,J2,J3; Bracket br[2]; bool nh; int nb=classify_mode(cv,J1,J2,J3,br,nh);
  printf("J1=%g J2=%g J3=%g nb=%d need_hump=%d\n",J1,J2,J3,nb,nh);
  // scan P1 and P2 near t=1
  for(int p=1;p<=2;p++){ printf("piece %d\n",p); for(int k=0;k<=30;k++){ double t=1.0-pow(10.0,-k/3.0)*(1-g.tmin); double d; double gg=curve_gd(cv,p,t,d); printf("  1-t=%.3e g=%.6g dg=%.4g\n",1-t,gg,d);} }
  PairGeom g0=g; printf("range_max=%g\n", range_max(ice,g0));
}
