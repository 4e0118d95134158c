// mini_eigen: matrix exponential by scaling and squaring with a Pade approximant (Higham 2005), the
// algorithm behind MatrixBase::exp() of unsupported/Eigen/MatrixFunctions.  Degree thresholds for
// single precision: (3) 4.258730016922831e-1, (5) 1.880152677804762, (7) 3.925724783138660 with
// squarings; double precision adds degrees 9 and 13.  TEST INFRASTRUCTURE ONLY.
#ifndef MINI_EIGEN_EXPM_H
#define MINI_EIGEN_EXPM_H
#include <cstdlib>
namespace Eigen {
namespace mini_internal {

template <class M>
void pade(int degree, const M& A, M& U, M& V) {
    typedef typename M::Scalar S;
    static const double b3[] = {120., 60., 12., 1.};
    static const double b5[] = {30240., 15120., 3360., 420., 30., 1.};
    static const double b7[] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
    static const double b9[] = {17643225600., 8821612800., 2075673600., 302702400., 30270240.,
                                2162160., 110880., 3960., 90., 1.};
    const double* b = degree == 3 ? b3 : degree == 5 ? b5 : degree == 7 ? b7 : b9;
    const int n = A.rows();
    M I(n, n);
    I.setIdentity();
    M A2 = A * A;
    M tmp = S(b[3]) * A2 + S(b[1]) * I;
    V = S(b[2]) * A2 + S(b[0]) * I;
    if (degree >= 5) {
        M A4 = A2 * A2;
        tmp = S(b[5]) * A4 + tmp;
        V = S(b[4]) * A4 + V;
        if (degree >= 7) {
            M A6 = A4 * A2;
            tmp = S(b[7]) * A6 + tmp;
            V = S(b[6]) * A6 + V;
            if (degree >= 9) {
                M A8 = A6 * A2;
                tmp = S(b[9]) * A8 + tmp;
                V = S(b[8]) * A8 + V;
            }
        }
    }
    U = A * tmp;
}

template <class M>
void pade13(const M& A, M& U, M& V) {
    typedef typename M::Scalar S;
    static const double b[] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                               129060195264000.,   10559470521600.,    670442572800.,     33522128640.,
                               1323241920.,        40840800.,          960960.,           16380.,
                               182.,               1.};
    const int n = A.rows();
    M I(n, n);
    I.setIdentity();
    M A2 = A * A, A4 = A2 * A2, A6 = A4 * A2;
    V = S(b[13]) * A6 + S(b[11]) * A4 + S(b[9]) * A2;
    M tmp = A6 * V;
    tmp += S(b[7]) * A6 + S(b[5]) * A4 + S(b[3]) * A2 + S(b[1]) * I;
    U = A * tmp;
    tmp = S(b[12]) * A6 + S(b[10]) * A4 + S(b[8]) * A2;
    V = A6 * tmp;
    V += S(b[6]) * A6 + S(b[4]) * A4 + S(b[2]) * A2 + S(b[0]) * I;
}

}   // namespace mini_internal

template <class D>
typename MatrixBase<D>::PlainObject MatrixBase<D>::exp() const {
    typedef typename MatrixBase<D>::PlainObject M;
    assert(rows() == cols());
    const int n = rows();
    M A = eval(), U(n, n), V(n, n);
    Scalar l1 = 0;
    for (int j = 0; j < n; ++j) {
        Scalar s = 0;
        for (int i = 0; i < n; ++i) s += std::fabs(A.coeff(i, j));
        if (s > l1) l1 = s;
    }
    int squarings = 0;
    // Test hook: with MINI_EIGEN_EXP_NILPOTENT3 set, a matrix with A*A*A == 0 is exponentiated by its
    // finite series I + A + A*A/2 (what the oracle's restatement evaluates), so that everything
    // AROUND the exponential can be compared bit for bit.  Default is the Pade evaluation below.
    if (std::getenv("MINI_EIGEN_EXP_NILPOTENT3")) {
        M A2 = A * A, A3 = A2 * A;
        bool nil = true;
        for (int j = 0; j < n && nil; ++j)
            for (int i = 0; i < n; ++i)
                if (A3.coeff(i, j) != Scalar(0)) { nil = false; break; }
        if (nil) {
            M E(n, n);
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) {
                    Scalar e = Scalar(i == j ? 1 : 0) + A.coeff(i, j);
                    E.coeffRef(i, j) = e + Scalar(0.5) * A2.coeff(i, j);
                }
            return E;
        }
    }
    if (sizeof(Scalar) == sizeof(float)) {
        if (l1 < Scalar(4.258730016922831e-001)) mini_internal::pade(3, A, U, V);
        else if (l1 < Scalar(1.880152677804762e+000)) mini_internal::pade(5, A, U, V);
        else {
            std::frexp(l1 / Scalar(3.925724783138660), &squarings);
            if (squarings < 0) squarings = 0;
            A = A * Scalar(std::ldexp(1.0, -squarings));
            mini_internal::pade(7, A, U, V);
        }
    } else {
        if (l1 < Scalar(1.495585217958292e-002)) mini_internal::pade(3, A, U, V);
        else if (l1 < Scalar(2.539398330063230e-001)) mini_internal::pade(5, A, U, V);
        else if (l1 < Scalar(9.504178996162932e-001)) mini_internal::pade(7, A, U, V);
        else if (l1 < Scalar(2.097847961257068e+000)) mini_internal::pade(9, A, U, V);
        else {
            std::frexp(l1 / Scalar(5.371920351148152), &squarings);
            if (squarings < 0) squarings = 0;
            A = A * Scalar(std::ldexp(1.0, -squarings));
            mini_internal::pade13(A, U, V);
        }
    }
    M numer = U + V, denom = V - U;
    mini_lu_solve<Scalar>(n, n, denom.data(), numer.data());
    for (int s = 0; s < squarings; ++s) numer = numer * numer;
    return numer;
}
}   // namespace Eigen
#endif
