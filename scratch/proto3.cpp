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
// safeguarded Newton in [a,b] (ga, gb opposite signs), start x0 (NaN: secant point)
static double newton(const IceParams&ice, const PairGeom&g, const PC&pc, bool turned, double a, double ga, double b, double gb, double x0, int&iters){
  double x = x0;
  if(!(x > fmin(a,b) && x < fmax(a,b))) x = (a*gb - b*ga)/(gb-ga);
  if(!(x > fmin(a,b) && x < fmax(a,b))) x = 0.5*(a+b);
  iters=0;
  for(int it=0; it<100; ++it){
    double dg; double gx = evalgd(ice,g,pc,turned,x,dg); ++iters;
    if(fabs(gx) <= 1e-10) break;
    if((gx>0)==(gb>0)) { b=x; gb=gx; } else { a=x; ga=gx; }
    double xn = x - gx/dg;
    double lo=fmin(a,b), hi=fmax(a,b);
    if(!(xn>lo && xn<hi)) { // fall back: secant on the bracket, else bisection
      xn = 0.5*(a+b);
    } else if (fabs(xn-x) <= 1e-9*(fabs(x)+1e-3)) { x = xn; break; }
    if(hi-lo <= 4e-16*(fabs(lo)+fabs(hi))) { x = xn; break; }
    x = xn;
  }
  return x;
}
int main(int argc,char**argv){
  IceParams ice; ice.n_ice=1.78; ice.dn=0.423; ice.z0=77; ice.inv_z0=1/77.; ice.ns=ice.n_ice-ice.dn; ice.n_refl=0; ice.zr=-1e30; ice.gr=0; ice.nr=ice.n_ice; ice.att_model=0;
  int mode = argc>1?atoi(argv[1]):0;
  std::mt19937_64 rng(5); std::uniform_real_distribution<double> U(0,1);
  int N=200000; long cnt[4]={0}, ev[4]={0}, hist[4][40]={{0}}; double maxdb[4]={0};
  for(int i=0;i<N;i++){
    double r=sqrt(U(rng))*6000, ph=U(rng)*2*M_PI, z=-2700*U(rng);
    int st=(int)(U(rng)*25); double ax=((st%5)-2)*1500., ay=((st/5)-2)*1500., az=-145-5*(int)(U(rng)*4);
    Frame2D f; make_frame(r*cos(ph), r*sin(ph), z, ax, ay, az, f);
    PairGeom g; make_pair_geom(ice, f.z1, f.z2, fmax(f.rho,1e-12), g);
    Curve cv; cv.ice=&ice; cv.g=&g; cv.m_dir=mode_coeffs(0,1,false); cv.m_trn=mode_coeffs(0,1,true);
    PC sub{ice.ns, g.c0_sub, g.A1, g.A2, false}, band{g.n2, g.c0_band, g.B1, 0.0, true};
    double tmin = ice.ns/(g.n2+g.s2max);
    double J[5]; J[0]=J[4]=-g.rho; J[1]=curve_g(cv,0,1.0); J[3]=curve_g(cv,3,1.0); J[2]=curve_g(cv,1,0.0);
    double pa_old[4]={0,g.s2max,0,1}, pb_old[4]={1,0,g.s2max,0};
    double pa[4]={0,tmin,1,1}, pb[4]={1,1,tmin,0};
    for(int p=0;p<4;p++) if((J[p]>0)!=(J[p+1]>0)){
      double vold = solve_piece(cv,p,pa_old[p],J[p],pb_old[p],J[p+1]);
      RayState rs; ray_state(ice,g,(p==1||p==2),vold,rs);
      bool isb=(p==1||p==2); const PC&pc=isb?band:sub;
      double x0 = NAN;
      if(mode>=1 && p<2){ // direct: straight-line guess with path-averaged index
        double dz = g.z2-g.z1; double nbar = ice.n_ice - ice.z0*(g.g2-g.g1)/fmax(dz,1e-9);
        double sinth = g.rho/sqrt(g.rho*g.rho+dz*dz); double b0 = nbar*sinth;
        // beta -> t within class: beta = nX 2t/(1+t^2) -> t = beta/(nX + sqrt(nX^2-beta^2))
        if(b0 < pc.nX) x0 = b0/(pc.nX+sqrt(pc.nX*pc.nX-b0*b0));
      }
      if(mode>=2 && p==3){ // reflected: image straight line
        double dz = -g.z1-g.z2; double nbar = ice.n_ice - ice.z0*((ice.dn-g.g1)+(ice.dn-g.g2))/fmax(dz,1e-9);
        double sinth = g.rho/sqrt(g.rho*g.rho+dz*dz); double b0 = nbar*sinth;
        if(b0 < pc.nX) x0 = b0/(pc.nX+sqrt(pc.nX*pc.nX-b0*b0));
      }
      int iters; double t = newton(ice,g,pc,p>=2,pa[p],J[p],pb[p],J[p+1],x0,iters);
      double dg, beta; evalgd(ice,g,pc,p>=2,t,dg,&beta);
      double db = fabs(beta-rs.beta)/rs.beta; if(db>maxdb[p]) maxdb[p]=db;
      cnt[p]++; ev[p]+=iters; hist[p][iters>39?39:iters]++;
    }
  }
  for(int p=0;p<4;p++){ printf("piece %d: %ld roots avg %.2f gd-evals maxdbeta %.2e | ", p, cnt[p], cnt[p]?(double)ev[p]/cnt[p]:0., maxdb[p]); for(int k=0;k<40;k++) if(hist[p][k]) printf("%d:%ld ",k,hist[p][k]); printf("\n"); }
}
