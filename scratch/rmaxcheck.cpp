#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <vector>
#include <random>
#include "../nuradiomc_b200/csrc/nrmc_math.cuh"
using namespace nrmc;
int main(int argc,char**argv){
  double z0=77, dn=0.423; if(argc>1){ z0=atof(argv[1]); dn=atof(argv[2]); }
  IceParams ice; ice.n_ice=1.78; ice.dn=dn; ice.z0=z0; ice.inv_z0=1/z0; ice.ns=ice.n_ice-ice.dn; ice.n_refl=0; ice.zr=-1e30; ice.gr=0; ice.nr=ice.n_ice; ice.att_model=0;
  const int n=401; const double dz=8.0; std::vector<double> T(n*n);
  for(int i1=0;i1<n;i1++) for(int i2=0;i2<=i1;i2++){ PairGeom g; make_pair_geom(ice,-i1*dz,-i2*dz,0.0,g); double r=range_max(ice,g); if(!(r<1e6)) r=INFINITY; T[i1*n+i2]=T[i2*n+i1]=r; }
  for(int i1=0;i1<n;i1++) for(int i2=0;i2<n;i2++){ double v=T[i1*n+i2]; if(i1>0) v=fmax(v,T[(i1-1)*n+i2]); if(i2>0) v=fmax(v,T[i1*n+i2-1]); T[i1*n+i2]=v; }
  // monotonicity of the table itself
  long viol=0; for(int i1=1;i1<n;i1++) for(int i2=0;i2<n;i2++){ if(T[i1*n+i2] < T[(i1-1)*n+i2]*(1-1e-12)) viol++; }
  printf("table monotonicity violations: %ld ; T[0,0]=%g T[19,19]=%g T[400,19]=%g T[400,400]=%g\n", viol, T[0], T[19*n+19], T[400*n+19], T[400*n+400]);
  std::mt19937_64 rng(7); std::uniform_real_distribution<double> U(0,1);
  long shadow=0, rejected=0, wrong=0, N=400000; double worst=1e300;
  TraceOutputs o = {0};
  for(long i=0;i<N;i++){
    double z1=-3100*U(rng), z2 = (i%3==0)? -200*U(rng) : -3100*U(rng); if(z2<z1) std::swap(z1,z2);
    PairGeom g0; make_pair_geom(ice,z1,z2,1.0,g0); double rm=range_max(ice,g0);
    // rho around the horizon: stress the boundary
    double rho = (i%2)? rm*(0.9+0.4*U(rng)) : 8000*U(rng);
    PairGeom g; make_pair_geom(ice,z1,z2,fmax(rho,1e-12),g);
    Root roots[2]; int nr=find_roots_mode(ice,g,0,1,roots);
    int i1=(int)ceil(-z1/dz), i2=(int)ceil(-z2/dz);
    bool rej = (i1<n && i2<n) && rho > T[i1*n+i2]*(1+1e-9)+1e-6;
    if(nr==0) shadow++;
    if(rej){ rejected++; if(nr>0){ wrong++; if(wrong<5) printf("WRONG z1=%g z2=%g rho=%.6f bound=%.6f rm=%.6f\n",z1,z2,rho,T[i1*n+i2],rm);} }
    if(nr>0 && i1<n && i2<n){ double slack = T[i1*n+i2]-rho; if(slack<worst) worst=slack; }
  }
  printf("pairs %ld shadow %ld rejected-by-table %ld (%.1f%% of shadow) WRONG %ld ; smallest (bound - rho) among pairs with solutions: %g m\n", N, shadow, rejected, 100.0*rejected/shadow, wrong, worst);
}
