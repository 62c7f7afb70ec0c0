// prototype: uniform t-parametrisation, value+derivative in one evaluation, safeguarded Newton
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
static long g_evals = 0;
#define NRMC_COUNT_EVALS
#include "../nuradiomc_b200/csrc/nrmc_math.cuh"
#include <random>
using namespace nrmc;
static long n_gd = 0;
struct PC { double nX, c0, O1, O2; bool band; };
static inline double rsq(double x){ return 1.0/sqrt(x); }
// value and derivative (k = 0 mode)
static double evalgd(const IceParams&ice, const PairGeom&g, const PC&pc, bool turned, double t, double&dg, double *beta_out=nullptr){
  ++n_gd;
  double q = 1.0/(1.0+t*t);
  double beta = pc.nX*2.0*t*q, sig = pc.nX*(1.0-t)*(1.0+t)*q;
  double sg2 = sig*sig, c = pc.c0+sg2;
  double irc = rsq(c), rc = c*irc;
  double x1 = pc.O1+sg2, x2 = pc.O2+sg2;
  double is1 = rsq(fmax(x1,1e-300)), is2 = rsq(fmax(x2,1e-300));
  double s1 = x1*is1, s2 = x2*is2;
  double k1 = rc*s1 + (c - ice.n_ice*g.g1), k2 = rc*s2 + (c - ice.n_ice*g.g2);
  double bp = 2.0*sig*q, sp = -2.0*beta*q, h = sig*sp;
  double ds1 = sp*(sig*is1), ds2 = pc.band ? sp : sp*(sig*is2);
  double drc = h*irc, dc = 2.0*h;
  double dk1 = drc*s1 + rc*ds1 + dc, dk2 = drc*s2 + rc*ds2 + dc;
  double A = beta*irc, dA = (bp - A*drc)*irc;
  double P, dlnP, lin;
  if(!turned){ double iv = 1.0/(k1*k2); P = k2*k2*iv; dlnP = (dk2*k1 - dk1*k2)*iv; lin = g.z2 - g.z1; }
  else { double KT, dKT; if(!pc.band){ KT = rc*sig + (c - ice.n_ice*ice.dn); dKT = drc*sig + rc*sp + dc; } else { KT = ice.dn*beta; dKT = ice.dn*bp; }
    double iv = 1.0/(k1*k2*KT); P = KT*KT*KT*iv; dlnP = (2.0*dKT*k1*k2 - dk1*k2*KT - dk2*k1*KT)*iv; lin = -g.z1-g.z2; }
  double Bk = lin - ice.z0*log(P);
  double R = A*Bk; dg = dA*Bk - A*ice.z0*dlnP;
  if(beta_out) *beta_out = beta;
  return R - g.rho;
}
int main(int argc,char**argv){
  IceParams ice; ice.n_ice=1.78; ice.dn=0.423; ice.z0=77; ice.inv_z0=1/77.; ice.ns=ice.n_ice-ice.dn; ice.n_refl=0; ice.zr=-1e30; ice.gr=0; ice.nr=ice.n_ice; ice.att_model=0;
  std::mt19937_64 rng(5); std::uniform_real_distribution<double> U(0,1);
  int N=100000; double maxerr=0, maxderr=0;
  for(int i=0;i<N;i++){
    double r=sqrt(U(rng))*6000, ph=U(rng)*2*M_PI, z=-2700*U(rng);
    int st=(int)(U(rng)*25); double ax=((st%5)-2)*1500., ay=((st/5)-2)*1500., az=-145-5*(int)(U(rng)*4);
    Frame2D f; make_frame(r*cos(ph), r*sin(ph), z, ax, ay, az, f);
    PairGeom g; make_pair_geom(ice, f.z1, f.z2, fmax(f.rho,1e-12), g);
    Curve cv; cv.ice=&ice; cv.g=&g; cv.m_dir=mode_coeffs(0,1,false); cv.m_trn=mode_coeffs(0,1,true);
    PC sub{ice.ns, g.c0_sub, g.A1, g.A2, false}, band{g.n2, g.c0_band, g.B1, 0.0, true};
    double tmin = ice.ns/(g.n2+g.s2max);
    // compare values: sub pieces at random t, band at random t in [tmin,1]
    for(int p=0;p<4;p++){
      bool isb = (p==1||p==2); double t = isb ? tmin + (1-tmin)*U(rng) : U(rng);
      const PC&pc = isb?band:sub; double dg, beta;
      double gn = evalgd(ice,g,pc,p>=2,t,dg,&beta);
      double vold = isb ? g.n2*(1-t*t)/(1+t*t) : t;
      double go = curve_g(cv,p,vold);
      double e = fabs(gn-go)/(fabs(go)+g.rho); if(e>maxerr){maxerr=e; }
      // derivative check by central difference
      double hh=1e-6, d1,d2; double gp=evalgd(ice,g,pc,p>=2,t+hh,d1), gm=evalgd(ice,g,pc,p>=2,t-hh,d2);
      double fd=(gp-gm)/(2*hh); double de=fabs(fd-dg)/(fabs(dg)+1e-3*fabs(go+g.rho)+1e-9); if(de>maxderr && t>1e-3 && t<1-1e-3 && t>tmin+1e-3){maxderr=de; if(de>1e-4) printf("p%d t=%g dg=%g fd=%g g=%g\n",p,t,dg,fd,gn);}
    }
  }
  printf("max rel value err %.3e, max rel deriv err %.3e\n", maxerr, maxderr);
}
