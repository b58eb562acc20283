// d(rho)/dT and d(rho)/dS of the five equations of state Veros supports
// (veros/core/density/get_rho.py:93-131).  p is |zt[k]| (isoneutral.py:40-41).
//   1 linear            linear_eq.py:37-43
//   2 Vallis, no p      nonlinear_eq1.py:40-48
//   3 Vallis with p     nonlinear_eq2.py:58-66
//   4 Vallis, no S      nonlinear_eq3.py:34-41
//   5 TEOS-10 48-term   gsw.py:105-274
#pragma once

namespace vb {

template <int EOS>
struct Eos {
    // cheap types are recomputed wherever needed; type 5 is evaluated once per cell into a workspace
    static constexpr bool kExpensive = (EOS == 5);
};

__device__ __forceinline__ void gsw_drho(double sa, double ct, double p, double& drdT, double& drdS) {
    constexpr double v01 = 9.998420897506056e2, v02 = 2.839940833161907e0, v03 = -3.147759265588511e-2,
                     v04 = 1.181805545074306e-3, v05 = -6.698001071123802e0, v06 = -2.986498947203215e-2,
                     v07 = 2.327859407479162e-4, v08 = -3.988822378968490e-2, v09 = 5.095422573880500e-4,
                     v10 = -1.426984671633621e-5, v11 = 1.645039373682922e-7, v12 = -2.233269627352527e-2,
                     v13 = -3.436090079851880e-4, v14 = 3.726050720345733e-6, v15 = -1.806789763745328e-4,
                     v16 = 6.876837219536232e-7, v17 = -3.087032500374211e-7, v18 = -1.988366587925593e-8,
                     v19 = -1.061519070296458e-11, v20 = 1.550932729220080e-10, v21 = 1.0e0,
                     v22 = 2.775927747785646e-3, v23 = -2.349607444135925e-5, v24 = 1.119513357486743e-6,
                     v25 = 6.743689325042773e-10, v26 = -7.521448093615448e-3, v27 = -2.764306979894411e-5,
                     v28 = 1.262937315098546e-7, v29 = 9.527875081696435e-10, v30 = -1.811147201949891e-11,
                     v31 = -3.303308871386421e-5, v32 = 3.801564588876298e-7, v33 = -7.672876869259043e-9,
                     v34 = -4.634182341116144e-11, v35 = 2.681097235569143e-12, v36 = 5.419326551148740e-6,
                     v37 = -2.742185394906099e-5, v38 = -3.212746477974189e-7, v39 = 3.191413910561627e-9,
                     v40 = -1.931012931541776e-12, v41 = -1.105097577149576e-7, v42 = 6.211426728363857e-10,
                     v43 = -1.119011592875110e-10, v44 = -1.941660213148725e-11, v45 = -1.864826425365600e-14,
                     v46 = 1.119522344879478e-14, v47 = -1.200507748551599e-15, v48 = 6.057902487546866e-17;
    constexpr double a01 = 2.839940833161907e0, a02 = -6.295518531177023e-2, a03 = 3.545416635222918e-3,
                     a04 = -2.986498947203215e-2, a05 = 4.655718814958324e-4, a06 = 5.095422573880500e-4,
                     a07 = -2.853969343267241e-5, a08 = 4.935118121048767e-7, a09 = -3.436090079851880e-4,
                     a10 = 7.452101440691467e-6, a11 = 6.876837219536232e-7, a12 = -1.988366587925593e-8,
                     a13 = -2.123038140592916e-11, a14 = 2.775927747785646e-3, a15 = -4.699214888271850e-5,
                     a16 = 3.358540072460230e-6, a17 = 2.697475730017109e-9, a18 = -2.764306979894411e-5,
                     a19 = 2.525874630197091e-7, a20 = 2.858362524508931e-9, a21 = -7.244588807799565e-11,
                     a22 = 3.801564588876298e-7, a23 = -1.534575373851809e-8, a24 = -1.390254702334843e-10,
                     a25 = 1.072438894227657e-11, a26 = -3.212746477974189e-7, a27 = 6.382827821123254e-9,
                     a28 = -5.793038794625329e-12, a29 = 6.211426728363857e-10, a30 = -1.941660213148725e-11,
                     a31 = -3.729652850731201e-14, a32 = 1.119522344879478e-14, a33 = 6.057902487546866e-17;
    constexpr double b01 = -6.698001071123802e0, b02 = -2.986498947203215e-2, b03 = 2.327859407479162e-4,
                     b04 = -5.983233568452735e-2, b05 = 7.643133860820750e-4, b06 = -2.140477007450431e-5,
                     b07 = 2.467559060524383e-7, b08 = -1.806789763745328e-4, b09 = 6.876837219536232e-7,
                     b10 = 1.550932729220080e-10, b11 = -7.521448093615448e-3, b12 = -2.764306979894411e-5,
                     b13 = 1.262937315098546e-7, b14 = 9.527875081696435e-10, b15 = -1.811147201949891e-11,
                     b16 = -4.954963307079632e-5, b17 = 5.702346883314446e-7, b18 = -1.150931530388857e-8,
                     b19 = -6.951273511674217e-11, b20 = 4.021645853353715e-12, b21 = 1.083865310229748e-5,
                     b22 = -1.105097577149576e-7, b23 = 6.211426728363857e-10, b24 = 1.119522344879478e-14;

    const double sqrtsa = sqrt(sa);
    const double den = v01 + ct * (v02 + ct * (v03 + v04 * ct)) +
                       sa * (v05 + ct * (v06 + v07 * ct) + sqrtsa * (v08 + ct * (v09 + ct * (v10 + v11 * ct)))) +
                       p * (v12 + ct * (v13 + v14 * ct) + sa * (v15 + v16 * ct) +
                            p * (v17 + ct * (v18 + v19 * ct) + v20 * sa));
    const double num = v21 + ct * (v22 + ct * (v23 + ct * (v24 + v25 * ct))) +
                       sa * (v26 + ct * (v27 + ct * (v28 + ct * (v29 + v30 * ct))) + v36 * sa +
                             sqrtsa * (v31 + ct * (v32 + ct * (v33 + ct * (v34 + v35 * ct))))) +
                       p * (v37 + ct * (v38 + ct * (v39 + v40 * ct)) + sa * (v41 + v42 * ct) +
                            p * (v43 + ct * (v44 + v45 * ct + v46 * sa) + p * (v47 + v48 * ct)));
    const double rec_num = 1.0 / num;
    const double rho = rec_num * den;

    const double dden_dct = a01 + ct * (a02 + a03 * ct) +
                            sa * (a04 + a05 * ct + sqrtsa * (a06 + ct * (a07 + a08 * ct))) +
                            p * (a09 + a10 * ct + a11 * sa + p * (a12 + a13 * ct));
    const double dnum_dct = a14 + ct * (a15 + ct * (a16 + a17 * ct)) +
                            sa * (a18 + ct * (a19 + ct * (a20 + a21 * ct)) +
                                  sqrtsa * (a22 + ct * (a23 + ct * (a24 + a25 * ct)))) +
                            p * (a26 + ct * (a27 + a28 * ct) + a29 * sa + p * (a30 + a31 * ct + a32 * sa + a33 * p));
    drdT = (dden_dct - dnum_dct * rho) * rec_num;

    const double dden_dsa = b01 + ct * (b02 + b03 * ct) + sqrtsa * (b04 + ct * (b05 + ct * (b06 + b07 * ct))) +
                            p * (b08 + b09 * ct + b10 * p);
    const double dnum_dsa = b11 + ct * (b12 + ct * (b13 + ct * (b14 + b15 * ct))) +
                            sqrtsa * (b16 + ct * (b17 + ct * (b18 + ct * (b19 + b20 * ct)))) + b21 * sa +
                            p * (b22 + ct * (b23 + b24 * p));
    drdS = (dden_dsa - dnum_dsa * rho) * rec_num;
}

// drdT, drdS WITHOUT the maskT factor.
template <int EOS>
__device__ __forceinline__ void eos_drho(double sa, double ct, double p, double& drdT, double& drdS) {
    constexpr double rho0 = 1024.0, theta0 = 283.0 - 273.15, betaT = 1.67e-4, betaS = 0.78e-3, grav = 9.81;
    if (EOS == 1) {
        drdT = -betaT * rho0;
        drdS = betaS * rho0;
    } else if (EOS == 2 || EOS == 4) {
        constexpr double betaTs = 1e-5 / 2.0;
        const double thetas = ct - theta0;
        drdT = -(betaT + 2 * betaTs * thetas) * rho0;
        drdS = (EOS == 2) ? betaS * rho0 : 0.0;
    } else if (EOS == 3) {
        constexpr double betaTs = 1e-5, gammas = 1.1e-8;
        const double zz = -p;
        const double thetas = ct - theta0;
        drdT = -(betaTs * thetas + betaT * (1 - gammas * grav * zz * rho0)) * rho0;
        drdS = betaS * rho0;
    } else {
        gsw_drho(sa, ct, p, drdT, drdS);
    }
}

}  // namespace vb
