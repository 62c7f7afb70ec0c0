// quadrature-order study for the SP1 attenuation integral in the u-variable (CPU, test only)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <random>
#include "../nuradiomc_b200/csrc/nrmc_att.cuh"
#include "gl_tables.h"
using namespace nrmc;
static void integrate(const IceParams&ice, const AttPlan&plan, const double*X, const double*W, int nq, int spp_override, const double*lnf, const int*band, int F, double*I){
  Gl3Table gl3{nullptr,0};
  for(int j=0;j<F;j++) I[j]=0;
  for(int panel=0;panel<3;panel++){
    int mult = plan_mult(plan,0,panel);
    double plo = panel==0?plan.uT:(panel==1?plan.u2:plan.u1), phi = panel==0?plan.u2:(panel==1?plan.u1:plan.ur);
    if(mult<=0 || !(phi>plo)) continue;
    for(int sub=0;sub<spp_override;sub++){
      double w=(phi-plo)/spp_override, lo=plo+sub*w, hi=lo+w;
      for(int i=0;i<nq;i++){
        double z,wds; att_node_geometry(ice,plan,lo,hi,X[i],W[i],z,wds);
        AttNode nd; att_node(1,z,gl3,nd);
        for(int j=0;j<F;j++){ double e=exp(nd.p0+(band[j]?nd.p2:nd.p1)*lnf[j]); I[j]+=mult*wds*fmin(e,1.0); }
      }
    }
  }
}
static double RMAX=6000;
int main(int argc,char**argv){ if(argc>1) RMAX=atof(argv[1]);
  IceParams ice; ice.n_ice=1.78; ice.dn=0.423; ice.z0=77; ice.inv_z0=1/77.; ice.ns=ice.n_ice-ice.dn; ice.n_refl=0; ice.zr=-1e30; ice.gr=0; ice.nr=ice.n_ice; ice.att_model=1;
  const int F=37; double f[F], lnf[F]; int band[F];
  for(int j=0;j<25;j++) f[j]=0.0048923679060665366+ j*(1.2-0.0048923679060665366)/24; for(int j=0;j<12;j++) f[25+j]=1.2+0.005+j*(2.5-1.205)/11;
  for(int j=0;j<F;j++){ lnf[j]=log(f[j]); band[j]=f[j]>=1.0; }
  std::mt19937_64 rng(5); std::uniform_real_distribution<double> U(0,1);
  struct Cfg{const double*X;const double*W;int nq;int spp;const char*name;} cfgs[]={{GLX13,GLW13,13,1,"13x1"},{GLX14,GLW14,14,1,"14x1"},{GLX10,GLW10,10,1,"10x1"},{GLX12,GLW12,12,1,"12x1"},{GLX16,GLW16,16,1,"16x1"},{GLX8,GLW8,8,2,"8x2"},{GLX16,GLW16,16,2,"16x2"}};
  const int NC=7; double maxrel[NC][2]={{0}}, maxabs[NC][2]={{0}};
  int N=40000, nsol=0;
  for(int i=0;i<N;i++){
    double r=sqrt(U(rng))*RMAX, ph=U(rng)*2*M_PI, z=-2700*U(rng);
    int st=(int)(U(rng)*25); double ax=((st%5)-2)*1500., ay=((st/5)-2)*1500., az=-145-5*(int)(U(rng)*4);
    if(i%3==0){ az=-2; } if(i%7==0){ z=-50*U(rng); }
    Frame2D fr; make_frame(r*cos(ph), r*sin(ph), z, ax, ay, az, fr);
    PairGeom g; make_pair_geom(ice, fr.z1, fr.z2, fmax(fr.rho,1e-12), g);
    Root roots[2]; int nr=find_roots_mode(ice,g,0,1,roots);
    for(int s=0;s<nr;s++){
      RayState rs; ray_state(ice,g,(roots[s].piece==1||roots[s].piece==2),roots[s].v,rs);
      AttPlan plan; att_plan(ice,g,roots[s].piece,0,1,rs,plan);
      double Iref[F], I[F]; integrate(ice,plan,GLX48,GLW48,48,4,lnf,band,F,Iref);
      int turned = roots[s].piece>=2; nsol++;
      for(int c=0;c<NC;c++){ integrate(ice,plan,cfgs[c].X,cfgs[c].W,cfgs[c].nq,cfgs[c].spp,lnf,band,F,I);
        for(int j=0;j<F;j++){ double a=exp(-I[j]), b=exp(-Iref[j]); double ae=fabs(a-b); if(b>1e-3){ double re=ae/b; if(re>maxrel[c][turned]) maxrel[c][turned]=re; } if(ae>maxabs[c][turned]) maxabs[c][turned]=ae; } }
    }
  }
  printf("%d solutions\n", nsol);
  for(int c=0;c<NC;c++) printf("%-6s direct: rel %.2e abs %.2e | turned: rel %.2e abs %.2e\n", cfgs[c].name, maxrel[c][0], maxabs[c][0], maxrel[c][1], maxabs[c][1]);
}
