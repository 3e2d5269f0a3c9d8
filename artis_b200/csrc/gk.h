// Adaptive Gauss-Kronrod quadrature on [a, b]: the rule the reference integrates its rate coefficients with
// (integrator.h:48-64 -> gausskronrod.h:173-257, a restatement of Boost.Math's gauss_kronrod<double, N>::integrate:
// apply the (2n+1)-point Kronrod rule, estimate the error from the embedded n-point Gauss rule, bisect while the error
// exceeds both the local and the inherited absolute tolerance, at most `max_depth` levels deep).
// The operation order follows the reference exactly (centre, Gauss nodes, Kronrod-only nodes; estimate of an interval
// = estimate(left half) + estimate(right half)) so that results agree to rounding. The recursion is unrolled onto an
// explicit stack (device code). Node and weight tables: tools/gk_tables.py (computed from first principles with
// mpmath; non-negative abscissae in increasing order starting with 0).
#pragma once
#include "hd.h"

namespace ab {

template <int NPOINTS>
struct GKTables;

template <>
struct GKTables<31> {
  static constexpr int N = 16;
  static constexpr int NG = 8;
  AHD static double abscissa(const int i) {
    constexpr double v[16] = {
        0.0,
        0.1011420669187174990270742,
        0.2011940939974345223006283,
        0.29918000715316881216678,
        0.3941513470775633698972074,
        0.4850818636402396806936557,
        0.5709721726085388475372267,
        0.6509967412974169705337359,
        0.7244177313601700474161861,
        0.7904185014424659329676493,
        0.8482065834104272162006483,
        0.8972645323440819008825097,
        0.9372733924007059043077589,
        0.967739075679139134257348,
        0.9879925180204854284895657,
        0.9980022986933970602851728,
    };
    return v[i];
  }
  AHD static double weights(const int i) {
    constexpr double v[16] = {
        0.1013300070147915490173748,
        0.1007698455238755950449467,
        0.09917359872179195933239317,
        0.09664272698362367850517991,
        0.09312659817082532122548687,
        0.08856444305621177064727544,
        0.08308050282313302103828925,
        0.07684968075772037889443278,
        0.06985412131872825870952008,
        0.06200956780067064028513923,
        0.05348152469092808726534315,
        0.0445897513247648766082273,
        0.03534636079137584622203795,
        0.025460847326715320186874,
        0.01500794732931612253837476,
        0.005377479872923348987792051,
    };
    return v[i];
  }
  AHD static double gauss_weights(const int i) {
    constexpr double v[8] = {
        0.2025782419255612728806202,
        0.1984314853271115764561183,
        0.1861610000155622110268006,
        0.1662692058169939335532009,
        0.1395706779261543144478048,
        0.1071592204671719350118695,
        0.07036604748810812470926742,
        0.03075324199611726835462839,
    };
    return v[i];
  }
};

template <>
struct GKTables<61> {
  static constexpr int N = 31;
  static constexpr int NG = 15;
  AHD static double abscissa(const int i) {
    constexpr double v[31] = {
        0.0,
        0.05147184255531769583302521,
        0.1028069379667370301470968,
        0.1538699136085835469637947,
        0.2045251166823098914389577,
        0.2546369261678898464398051,
        0.3040732022736250773726771,
        0.3527047255308781134710372,
        0.4004012548303943925354762,
        0.4470337695380891767806099,
        0.4924804678617785749936931,
        0.5366241481420198992641698,
        0.5793452358263616917560249,
        0.6205261829892428611404776,
        0.6600610641266269613700537,
        0.6978504947933157969322924,
        0.7337900624532268047261711,
        0.7677774321048261949179773,
        0.7997278358218390830136689,
        0.8295657623827683974428981,
        0.8572052335460610989586585,
        0.8825605357920526815431165,
        0.9055733076999077985465226,
        0.9262000474292743258793243,
        0.9443744447485599794158313,
        0.960021864968307512216871,
        0.9731163225011262683746939,
        0.9836681232797472099700326,
        0.9916309968704045948586284,
        0.9968934840746495402716301,
        0.9994844100504906375713259,
    };
    return v[i];
  }
  AHD static double weights(const int i) {
    constexpr double v[31] = {
        0.05149472942945156755834043,
        0.05142612853745902593386288,
        0.05122154784925877217065628,
        0.05088179589874960649229747,
        0.05040592140278234684089309,
        0.04979568342707420635781157,
        0.04905543455502977888752817,
        0.04818586175708712914077949,
        0.04718554656929915394526148,
        0.04605923827100698811627174,
        0.04481480013316266319235555,
        0.04345253970135606931683173,
        0.04196981021516424614714754,
        0.04037453895153595911199528,
        0.03867894562472759295034865,
        0.03688236465182122922391107,
        0.03497933802806002413749967,
        0.03298144705748372603181419,
        0.03090725756238776247288425,
        0.02875404876504129284397879,
        0.02650995488233310161060171,
        0.02419116207808060136568637,
        0.02182803582160919229716749,
        0.01941414119394238117340895,
        0.01692088918905327262757229,
        0.01436972950704580481245143,
        0.0118230152534963417422329,
        0.009273279659517763428441147,
        0.006630703915931292173319826,
        0.003890461127099884051267202,
        0.001389013698677007624551591,
    };
    return v[i];
  }
  AHD static double gauss_weights(const int i) {
    constexpr double v[15] = {
        0.1028526528935588403412856,
        0.101762389748405504596429,
        0.09959342058679526706278028,
        0.09636873717464425963946863,
        0.09212252223778612871763271,
        0.08689978720108297980238753,
        0.08075589522942021535469494,
        0.07375597473770520626824385,
        0.06597422988218049512812852,
        0.05749315621761906648172169,
        0.04840267283059405290293814,
        0.03879919256962704959680194,
        0.02878470788332336934971918,
        0.01846646831109095914230213,
        0.007968192496166605615465883,
    };
    return v[i];
  }
};
// the rule on [-1, 1] applied to g(x) = f(scale * x + mean); returns the Kronrod sum, `error` = the reference's estimate
template <int NPOINTS, class F>
AHD double gk_rule(const F& f, const double scale, const double mean, double& error) {
  using Tab = GKTables<NPOINTS>;
  constexpr int gauss_order = (NPOINTS - 1) / 2;
  constexpr bool centre_is_gauss_node = (gauss_order & 1) != 0;
  constexpr int gauss_start = centre_is_gauss_node ? 2 : 1;
  constexpr int kronrod_start = centre_is_gauss_node ? 1 : 2;
  const double f_centre = f((scale * 0.) + mean);
  double kronrod_result = f_centre * Tab::weights(0);
  double gauss_result = 0.;
  if constexpr (centre_is_gauss_node) {
    gauss_result += f_centre * Tab::gauss_weights(0);
  }
#pragma unroll
  for (int i = gauss_start; i < Tab::N; i += 2) {
    const double fp = f((scale * Tab::abscissa(i)) + mean);
    const double fm = f((scale * -Tab::abscissa(i)) + mean);
    kronrod_result += (fp + fm) * Tab::weights(i);
    gauss_result += (fp + fm) * Tab::gauss_weights(i / 2);
  }
#pragma unroll
  for (int i = kronrod_start; i < Tab::N; i += 2) {
    const double fp = f((scale * Tab::abscissa(i)) + mean);
    const double fm = f((scale * -Tab::abscissa(i)) + mean);
    kronrod_result += (fp + fm) * Tab::weights(i);
  }
  constexpr double eps = 2.220446049250313e-16;
  error = dmax(fabs(kronrod_result - gauss_result), fabs(kronrod_result * eps * 2));
  return kronrod_result;
}

// integral of f over [a, b] to the relative tolerance `tol` (gausskronrod.h:209-257 with max_depth = 15,
// integrator.h:61); `evals` (optional) counts applications of the rule
template <int NPOINTS, class F>
AHD double gk_integrate(const F& f, double a, double b, const double tol, int* evals = nullptr) {
  constexpr int MAX_DEPTH = 15;
  if (a == b) {
    return 0.;
  }
  double sign = 1.;
  if (b < a) {
    const double tmp = a;
    a = b;
    b = tmp;
    sign = -1.;
  }
  struct Frame {
    double a, b, abs_tol, left;
    int levels, state;
  };
  Frame stack[MAX_DEPTH + 1];
  int sp = 0;
  stack[0] = {a, b, 0., 0., MAX_DEPTH, 0};
  double ret = 0.;
  while (sp >= 0) {
    Frame& fr = stack[sp];
    if (fr.state == 0) {
      double error_local = 0.;
      const double mean = (fr.b + fr.a) / 2;
      const double scale = (fr.b - fr.a) / 2;
      const double r1 = gk_rule<NPOINTS>(f, scale, mean, error_local);
      if (evals != nullptr) {
        *evals += 1;
      }
      const double estimate = scale * r1;
      const double abs_tol1 = fabs(estimate * tol);
      if (fr.abs_tol == 0) {
        fr.abs_tol = abs_tol1;
      }
      if ((fr.levels != 0) && (abs_tol1 < error_local) && (fr.abs_tol < error_local)) {
        const double mid = (fr.a + fr.b) / 2;
        fr.state = 1;
        stack[sp + 1] = {fr.a, mid, fr.abs_tol / 2, 0., fr.levels - 1, 0};
        sp++;
        continue;
      }
      ret = estimate;
      sp--;
    } else if (fr.state == 1) {
      fr.left = ret;
      fr.state = 2;
      const double mid = (fr.a + fr.b) / 2;
      stack[sp + 1] = {mid, fr.b, fr.abs_tol / 2, 0., fr.levels - 1, 0};
      sp++;
    } else {
      ret = fr.left + ret;
      sp--;
    }
  }
  return (sign < 0.) ? -ret : ret;
}

}  // namespace ab
