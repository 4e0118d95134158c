// mini_eigen -- TEST INFRASTRUCTURE ONLY.
//
// Eigen is not installed in this image and the reference does not vendor it, so the reference's own
// quadruped/src/controllers/mpc/qr_mpc_interface.cpp cannot be compiled against the real library.
// This header provides, under the Eigen names, the small dense-matrix subset that file (and the
// reference headers it pulls in: utils/qr_cpptypes.h, qr_algebra.h, qr_se3.h, qr_print.hpp) needs,
// so that it compiles from /root/reference UNMODIFIED (oracle/Makefile target `refmpc`).
//
// It is not Eigen: there are no expression templates, every operator returns an evaluated matrix,
// and all products are plain sequential sums over the inner index in ascending order.  What a build
// against it pins is the reference's LOGIC (indexing, weights, block placement, solver set-up); the
// float32 rounding of a true Eigen build (vectorised GEMM, its own evaluation order) stays unpinned.
#ifndef MINI_EIGEN_CORE_H
#define MINI_EIGEN_CORE_H

#include <cassert>
#include <cmath>
#include <cstring>
#include <memory>
#include <ostream>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_STL_VECTOR_SPECIALIZATION_DEFINED

namespace Eigen {

const int Dynamic = -1;
enum NoChange_t { NoChange };
enum { ComputeFullU = 0x04, ComputeThinU = 0x08, ComputeFullV = 0x10, ComputeThinV = 0x20 };
enum TransformTraits { Isometry = 0x1, Affine = 0x2, AffineCompact = 0x12, Projective = 0x20 };
template <class T>
using aligned_allocator = std::allocator<T>;

template <class Scalar_, int Rows_, int Cols_>
class Matrix;
template <class M, int BR, int BC>
class Block;
template <class M>
class DiagonalView;
template <class T>
class Quaternion;
template <class M>
class JacobiSVD;
template <class T, int Dim, int Mode>
class Transform;

template <class Derived>
struct traits;
template <class S, int R, int C>
struct traits<Matrix<S, R, C>> {
    typedef S Scalar;
    enum { Rows = R, Cols = C };
};
template <class M, int BR, int BC>
struct traits<Block<M, BR, BC>> {
    typedef typename traits<M>::Scalar Scalar;
    enum { Rows = BR, Cols = BC };
};
template <class T>
struct traits<const T> : traits<T> {};
template <class M>
struct traits<DiagonalView<M>> {
    typedef typename traits<M>::Scalar Scalar;
    enum { Rows = Dynamic, Cols = 1 };
};

template <class XprType>
class CommaInitializer;
template <class M>
class ColwiseProxy;
template <class M>
class RowwiseProxy;
template <class M>
class QrSolveProxy;

// ------------------------------------------------------------------------------------------------
// Read-side interface shared by matrices and views.
template <class Derived>
class MatrixBase {
public:
    typedef typename traits<Derived>::Scalar Scalar;
    enum { RowsAtCompileTime = traits<Derived>::Rows, ColsAtCompileTime = traits<Derived>::Cols };
    typedef Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> PlainObject;
    typedef Matrix<Scalar, ColsAtCompileTime, RowsAtCompileTime> TransposedObject;

    const Derived& derived() const { return *static_cast<const Derived*>(this); }
    Derived& derived() { return *static_cast<Derived*>(this); }

    int rows() const { return derived().rows(); }
    int cols() const { return derived().cols(); }
    int size() const { return rows() * cols(); }
    Scalar coeff(int i, int j) const { return derived().coeff(i, j); }
    Scalar coeff(int i) const { return cols() == 1 ? coeff(i, 0) : coeff(0, i); }
    Scalar& coeffRef(int i, int j) { return derived().coeffRef(i, j); }
    Scalar& coeffRef(int i) { return cols() == 1 ? coeffRef(i, 0) : coeffRef(0, i); }
    Scalar operator()(int i, int j) const { return coeff(i, j); }
    Scalar& operator()(int i, int j) { return coeffRef(i, j); }
    Scalar operator()(int i) const { return coeff(i); }
    Scalar& operator()(int i) { return coeffRef(i); }
    Scalar operator[](int i) const { return coeff(i); }
    Scalar& operator[](int i) { return coeffRef(i); }

    PlainObject eval() const {
        PlainObject out(rows(), cols());
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) out.coeffRef(i, j) = coeff(i, j);
        return out;
    }
    TransposedObject transpose() const {
        TransposedObject out(cols(), rows());
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) out.coeffRef(j, i) = coeff(i, j);
        return out;
    }
    void transposeInPlace() {
        PlainObject t = eval();
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) coeffRef(i, j) = t.coeff(j, i);
    }

    // Views.
    Block<Derived, Dynamic, Dynamic> block(int i, int j, int r, int c) {
        return Block<Derived, Dynamic, Dynamic>(derived(), i, j, r, c);
    }
    Block<const Derived, Dynamic, Dynamic> block(int i, int j, int r, int c) const {
        return Block<const Derived, Dynamic, Dynamic>(derived(), i, j, r, c);
    }
    Block<Derived, RowsAtCompileTime, 1> col(int j) {
        return Block<Derived, RowsAtCompileTime, 1>(derived(), 0, j, rows(), 1);
    }
    Block<const Derived, RowsAtCompileTime, 1> col(int j) const {
        return Block<const Derived, RowsAtCompileTime, 1>(derived(), 0, j, rows(), 1);
    }
    Block<Derived, 1, ColsAtCompileTime> row(int i) {
        return Block<Derived, 1, ColsAtCompileTime>(derived(), i, 0, 1, cols());
    }
    Block<const Derived, 1, ColsAtCompileTime> row(int i) const {
        return Block<const Derived, 1, ColsAtCompileTime>(derived(), i, 0, 1, cols());
    }
    Block<Derived, Dynamic, 1> head(int n) { return Block<Derived, Dynamic, 1>(derived(), 0, 0, n, 1); }
    Block<const Derived, Dynamic, 1> head(int n) const {
        return Block<const Derived, Dynamic, 1>(derived(), 0, 0, n, 1);
    }
    Block<Derived, Dynamic, 1> tail(int n) { return Block<Derived, Dynamic, 1>(derived(), rows() - n, 0, n, 1); }
    Block<const Derived, Dynamic, 1> tail(int n) const {
        return Block<const Derived, Dynamic, 1>(derived(), rows() - n, 0, n, 1);
    }
    Block<Derived, Dynamic, 1> segment(int s, int n) { return Block<Derived, Dynamic, 1>(derived(), s, 0, n, 1); }
    DiagonalView<Derived> diagonal() { return DiagonalView<Derived>(derived()); }
    Block<Derived, Dynamic, Dynamic> topRows(int n) { return block(0, 0, n, cols()); }
    Block<Derived, Dynamic, Dynamic> bottomRows(int n) { return block(rows() - n, 0, n, cols()); }
    Block<Derived, Dynamic, Dynamic> leftCols(int n) { return block(0, 0, rows(), n); }
    Block<Derived, Dynamic, Dynamic> rightCols(int n) { return block(0, cols() - n, rows(), n); }
    Block<const Derived, Dynamic, Dynamic> topRows(int n) const { return block(0, 0, n, cols()); }
    Block<const Derived, Dynamic, Dynamic> bottomRows(int n) const { return block(rows() - n, 0, n, cols()); }
    Block<const Derived, Dynamic, Dynamic> leftCols(int n) const { return block(0, 0, rows(), n); }
    Block<const Derived, Dynamic, Dynamic> rightCols(int n) const { return block(0, cols() - n, rows(), n); }

    // Fixed-size views (`m.template block<3, 3>(i, j)` and friends).
#define MINI_EIGEN_VIEW2(NAME, I0, J0)                                                                      \
    template <int BR, int BC>                                                                               \
    Block<Derived, BR, BC> NAME() { return Block<Derived, BR, BC>(derived(), I0, J0, BR, BC); }             \
    template <int BR, int BC>                                                                               \
    Block<const Derived, BR, BC> NAME() const { return Block<const Derived, BR, BC>(derived(), I0, J0, BR, BC); }
    MINI_EIGEN_VIEW2(topLeftCorner, 0, 0)
    MINI_EIGEN_VIEW2(topRightCorner, 0, cols() - BC)
    MINI_EIGEN_VIEW2(bottomLeftCorner, rows() - BR, 0)
    MINI_EIGEN_VIEW2(bottomRightCorner, rows() - BR, cols() - BC)
#undef MINI_EIGEN_VIEW2
    template <int BR, int BC>
    Block<Derived, BR, BC> block(int i, int j) { return Block<Derived, BR, BC>(derived(), i, j, BR, BC); }
    template <int BR, int BC>
    Block<const Derived, BR, BC> block(int i, int j) const { return Block<const Derived, BR, BC>(derived(), i, j, BR, BC); }
    template <int N>
    Block<Derived, N, 1> head() { return Block<Derived, N, 1>(derived(), 0, 0, N, 1); }
    template <int N>
    Block<const Derived, N, 1> head() const { return Block<const Derived, N, 1>(derived(), 0, 0, N, 1); }
    template <int N>
    Block<Derived, N, 1> tail() { return Block<Derived, N, 1>(derived(), rows() - N, 0, N, 1); }
    template <int N>
    Block<const Derived, N, 1> tail() const { return Block<const Derived, N, 1>(derived(), rows() - N, 0, N, 1); }
    template <int N>
    Block<Derived, N, 1> segment(int s) { return Block<Derived, N, 1>(derived(), s, 0, N, 1); }
    template <int N>
    Block<const Derived, N, 1> segment(int s) const { return Block<const Derived, N, 1>(derived(), s, 0, N, 1); }
    template <int N>
    Block<Derived, N, ColsAtCompileTime> topRows() { return Block<Derived, N, ColsAtCompileTime>(derived(), 0, 0, N, cols()); }
    template <int N>
    Block<const Derived, N, ColsAtCompileTime> topRows() const { return Block<const Derived, N, ColsAtCompileTime>(derived(), 0, 0, N, cols()); }
    template <int N>
    Block<Derived, N, ColsAtCompileTime> bottomRows() { return Block<Derived, N, ColsAtCompileTime>(derived(), rows() - N, 0, N, cols()); }
    template <int N>
    Block<const Derived, N, ColsAtCompileTime> bottomRows() const { return Block<const Derived, N, ColsAtCompileTime>(derived(), rows() - N, 0, N, cols()); }
    template <int N>
    Block<Derived, RowsAtCompileTime, N> leftCols() { return Block<Derived, RowsAtCompileTime, N>(derived(), 0, 0, rows(), N); }
    template <int N>
    Block<const Derived, RowsAtCompileTime, N> leftCols() const { return Block<const Derived, RowsAtCompileTime, N>(derived(), 0, 0, rows(), N); }
    template <int N>
    Block<Derived, RowsAtCompileTime, N> rightCols() { return Block<Derived, RowsAtCompileTime, N>(derived(), 0, cols() - N, rows(), N); }
    template <int N>
    Block<const Derived, RowsAtCompileTime, N> rightCols() const { return Block<const Derived, RowsAtCompileTime, N>(derived(), 0, cols() - N, rows(), N); }
    Block<const Derived, Dynamic, 1> segment(int s, int n) const { return Block<const Derived, Dynamic, 1>(derived(), s, 0, n, 1); }
    TransposedObject adjoint() const { return transpose(); }

    // Reductions.
    Scalar trace() const {
        Scalar s = 0;
        for (int i = 0; i < rows(); ++i) s += coeff(i, i);
        return s;
    }
    Scalar sum() const {
        Scalar s = 0;
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) s += coeff(i, j);
        return s;
    }
    Scalar squaredNorm() const {
        Scalar s = 0;
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) s += coeff(i, j) * coeff(i, j);
        return s;
    }
    Scalar norm() const { return std::sqrt(squaredNorm()); }
    template <class O>
    Scalar dot(const MatrixBase<O>& o) const {
        Scalar s = 0;
        for (int i = 0; i < size(); ++i) s += coeff(i) * o.coeff(i);
        return s;
    }
    template <class O>
    Matrix<Scalar, 3, 1> cross(const MatrixBase<O>& o) const {
        Matrix<Scalar, 3, 1> r;
        r(0) = coeff(1) * o.coeff(2) - coeff(2) * o.coeff(1);
        r(1) = coeff(2) * o.coeff(0) - coeff(0) * o.coeff(2);
        r(2) = coeff(0) * o.coeff(1) - coeff(1) * o.coeff(0);
        return r;
    }

    // Setters (valid on anything writable).
    Derived& setZero() {
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) coeffRef(i, j) = Scalar(0);
        return derived();
    }
    Derived& setIdentity() {
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) coeffRef(i, j) = Scalar(i == j ? 1 : 0);
        return derived();
    }
    Derived& setConstant(Scalar v) {
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) coeffRef(i, j) = v;
        return derived();
    }
    Derived& noalias() { return derived(); }

    template <class O>
    Derived& operator+=(const MatrixBase<O>& o) {
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) coeffRef(i, j) += o.coeff(i, j);
        return derived();
    }
    template <class O>
    Derived& operator-=(const MatrixBase<O>& o) {
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) coeffRef(i, j) -= o.coeff(i, j);
        return derived();
    }
    Derived& operator*=(Scalar s) {
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) coeffRef(i, j) *= s;
        return derived();
    }
    Derived& operator/=(Scalar s) {
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) coeffRef(i, j) /= s;
        return derived();
    }

    PlainObject operator-() const {
        PlainObject out(rows(), cols());
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) out.coeffRef(i, j) = -coeff(i, j);
        return out;
    }

    CommaInitializer<Derived> operator<<(Scalar s);
    template <class O>
    CommaInitializer<Derived> operator<<(const MatrixBase<O>& o);

    // Coefficient-wise helpers, casts and broadcasting views used by the controller sources (robots/qr_robot.cpp,
    // planner/qr_foothold_planner.cpp, gait/qr_openloop_gait_generator.cpp, controllers/mpc/qr_mpc_stance_leg_controller.cpp).
    Scalar x() const { return coeff(0); }
    Scalar y() const { return coeff(1); }
    Scalar z() const { return coeff(2); }
    Scalar& x() { return coeffRef(0); }
    Scalar& y() { return coeffRef(1); }
    Scalar& z() { return coeffRef(2); }
    Derived& setOnes() { return setConstant(Scalar(1)); }
    template <class T>
    Matrix<T, RowsAtCompileTime, ColsAtCompileTime> cast() const {
        Matrix<T, RowsAtCompileTime, ColsAtCompileTime> out(rows(), cols());
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) out.coeffRef(i, j) = T(coeff(i, j));
        return out;
    }
    template <class O>
    PlainObject cwiseProduct(const MatrixBase<O>& o) const {
        PlainObject out(rows(), cols());
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) out.coeffRef(i, j) = coeff(i, j) * o.coeff(i, j);
        return out;
    }
    template <class O>
    PlainObject cwiseQuotient(const MatrixBase<O>& o) const {
        PlainObject out(rows(), cols());
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) out.coeffRef(i, j) = coeff(i, j) / o.coeff(i, j);
        return out;
    }
    ColwiseProxy<const Derived> colwise() const { return ColwiseProxy<const Derived>(derived()); }
    RowwiseProxy<const Derived> rowwise() const { return RowwiseProxy<const Derived>(derived()); }
    QrSolveProxy<Derived> colPivHouseholderQr() const { return QrSolveProxy<Derived>(derived()); }

    // Closed-form inverse of a 3x3 / general Gauss-Jordan otherwise (cofactors over the determinant).
    PlainObject inverse() const;
    // Scaling-and-squaring Pade exponential (unsupported/Eigen/MatrixFunctions), see mini_eigen_expm.h.
    PlainObject exp() const;
};

// ------------------------------------------------------------------------------------------------
// Storage: fixed sizes live in an in-object array, anything with a Dynamic extent on the heap.
template <class S, int R, int C, bool Fixed = (R != Dynamic && C != Dynamic)>
struct MatStorage;
template <class S, int R, int C>
struct MatStorage<S, R, C, true> {
    S d[R * C];
    MatStorage() { std::memset(d, 0, sizeof(d)); }
    void resize(int, int) {}
    int rows() const { return R; }
    int cols() const { return C; }
    S* data() { return d; }
    const S* data() const { return d; }
};
template <class S, int R, int C>
struct MatStorage<S, R, C, false> {
    std::vector<S> d;
    int r = (R == Dynamic ? 0 : R), c = (C == Dynamic ? 0 : C);
    void resize(int rr, int cc) {
        r = rr;
        c = cc;
        d.assign((size_t)rr * cc, S(0));
    }
    int rows() const { return r; }
    int cols() const { return c; }
    S* data() { return d.data(); }
    const S* data() const { return d.data(); }
};

// A 1 x 1 result (inner product written as a^T * b) converts to its scalar.
template <class Derived, class S, bool OneByOne>
struct ScalarConv {};
template <class Derived, class S>
struct ScalarConv<Derived, S, true> {
    operator S() const { return static_cast<const Derived*>(this)->coeff(0, 0); }
};

template <class Scalar_, int Rows_, int Cols_>
class Matrix : public MatrixBase<Matrix<Scalar_, Rows_, Cols_>>,
               public ScalarConv<Matrix<Scalar_, Rows_, Cols_>, Scalar_, Rows_ == 1 && Cols_ == 1> {
    MatStorage<Scalar_, Rows_, Cols_> st;

public:
    typedef Scalar_ Scalar;
    typedef MatrixBase<Matrix> Base;
    using Base::operator<<;

    Matrix() {}
    Matrix(int r, int c) { st.resize(r, c); }
    explicit Matrix(int n) {
        if (Rows_ == Dynamic || Cols_ == Dynamic) st.resize(Cols_ == 1 ? n : 1, Cols_ == 1 ? 1 : n);
    }
    Matrix(Scalar a, Scalar b, Scalar c) {
        st.resize(3, 1);
        st.data()[0] = a, st.data()[1] = b, st.data()[2] = c;
    }
    Matrix(Scalar a, Scalar b, Scalar c, Scalar d) {
        st.resize(4, 1);
        st.data()[0] = a, st.data()[1] = b, st.data()[2] = c, st.data()[3] = d;
    }
    template <class O>
    Matrix(const MatrixBase<O>& o) {
        assign(o);
    }
    Matrix(const Matrix&) = default;
    Matrix& operator=(const Matrix&) = default;
    template <class O>
    Matrix& operator=(const MatrixBase<O>& o) {
        assign(o);
        return *this;
    }
    template <class O>
    void assign(const MatrixBase<O>& o) {
        if (Rows_ == Dynamic || Cols_ == Dynamic) {
            if (rows() != o.rows() || cols() != o.cols()) st.resize(o.rows(), o.cols());
        } else if ((Rows_ == 1 || Cols_ == 1) && o.rows() == Cols_ && o.cols() == Rows_ && Rows_ != Cols_) {
            // Eigen transposes implicitly when a row vector is assigned to a column vector (and vice versa)
            for (int k = 0; k < Rows_ * Cols_; ++k) st.data()[k] = o.coeff(k);
            return;
        } else {
            assert(o.rows() == Rows_ && o.cols() == Cols_);
        }
        for (int j = 0; j < cols(); ++j)
            for (int i = 0; i < rows(); ++i) coeffRef(i, j) = o.coeff(i, j);
    }

    int rows() const { return st.rows(); }
    int cols() const { return st.cols(); }
    Scalar coeff(int i, int j) const { return st.data()[(size_t)j * rows() + i]; }
    Scalar& coeffRef(int i, int j) { return st.data()[(size_t)j * rows() + i]; }
    using Base::coeff;
    using Base::coeffRef;
    Scalar* data() { return st.data(); }
    const Scalar* data() const { return st.data(); }

    void resize(int r, int c) { st.resize(r, c); }
    void resize(int r, NoChange_t) { st.resize(r, cols()); }
    void resize(NoChange_t, int c) { st.resize(rows(), c); }
    void resize(int n) { st.resize(Cols_ == 1 ? n : 1, Cols_ == 1 ? 1 : n); }
    using Base::setZero;
    Matrix& setZero(int r, int c) {
        st.resize(r, c);
        return Base::setZero();
    }
    Matrix& setZero(int n) {
        resize(n);
        return Base::setZero();
    }
    void conservativeResize(int r, int c) {
        Matrix old = *this;
        st.resize(r, c);
        for (int j = 0; j < c && j < old.cols(); ++j)
            for (int i = 0; i < r && i < old.rows(); ++i) coeffRef(i, j) = old.coeff(i, j);
    }
    void conservativeResize(int n) { conservativeResize(Cols_ == 1 ? n : 1, Cols_ == 1 ? 1 : n); }

    static Matrix Zero() { return Matrix().setZero(); }
    static Matrix Zero(int r, int c) { return Matrix(r, c).setZero(); }
    static Matrix Zero(int n) { return Matrix(n).setZero(); }
    static Matrix Identity() { return Matrix().setIdentity(); }
    static Matrix Identity(int r, int c) { return Matrix(r, c).setIdentity(); }
    static Matrix Ones() { return Matrix().setConstant(Scalar(1)); }
    static Matrix Constant(Scalar v) { return Matrix().setConstant(v); }
    static Matrix Constant(int n, Scalar v) { return Matrix(n).setConstant(v); }
    static Matrix Constant(int r, int c, Scalar v) { return Matrix(r, c).setConstant(v); }
    template <class O>
    Matrix& operator*=(const MatrixBase<O>& o) {
        Matrix t = (*this) * o;
        *this = t;
        return *this;
    }
    using Base::operator*=;
};

typedef Matrix<float, Dynamic, Dynamic> MatrixXf;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<float, Dynamic, 1> VectorXf;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<double, 4, 1> Vector4d;

// ------------------------------------------------------------------------------------------------
// Rectangular view into another matrix (block / col / row / head / tail).
template <class M, int BR, int BC>
class Block : public MatrixBase<Block<M, BR, BC>> {
    M& m;
    int i0, j0, nr, nc;

public:
    typedef typename traits<M>::Scalar Scalar;
    typedef MatrixBase<Block> Base;
    using Base::operator<<;
    Block(M& mm, int i, int j, int r, int c) : m(mm), i0(i), j0(j), nr(r), nc(c) {
        assert(i >= 0 && j >= 0 && i + r <= mm.rows() && j + c <= mm.cols());
    }
    int rows() const { return nr; }
    int cols() const { return nc; }
    Scalar coeff(int i, int j) const { return m.coeff(i0 + i, j0 + j); }
    Scalar& coeffRef(int i, int j) { return const_cast<typename std::remove_const<M>::type&>(m).coeffRef(i0 + i, j0 + j); }
    using Base::coeff;
    using Base::coeffRef;
    template <class O>
    Block& operator=(const MatrixBase<O>& o) {
        typename MatrixBase<O>::PlainObject t = o.eval();   // the source may alias this view
        if ((nr == 1 || nc == 1) && o.rows() == nc && o.cols() == nr && nr != nc) {
            // Eigen transposes implicitly when a column vector is assigned to a row view (and vice versa)
            for (int k = 0; k < nr * nc; ++k) coeffRef(nc == 1 ? k : 0, nc == 1 ? 0 : k) = t.coeff(k);
            return *this;
        }
        assert(o.rows() == nr && o.cols() == nc);
        for (int j = 0; j < nc; ++j)
            for (int i = 0; i < nr; ++i) coeffRef(i, j) = t.coeff(i, j);
        return *this;
    }
    Block& operator=(const Block& o) { return operator=<Block>(o); }
};

template <class M>
class DiagonalView : public MatrixBase<DiagonalView<M>> {
    M& m;

public:
    typedef typename traits<M>::Scalar Scalar;
    typedef MatrixBase<DiagonalView> Base;
    using Base::operator<<;
    explicit DiagonalView(M& mm) : m(mm) {}
    int rows() const { return m.rows() < m.cols() ? m.rows() : m.cols(); }
    int cols() const { return 1; }
    Scalar coeff(int i, int) const { return m.coeff(i, i); }
    Scalar& coeffRef(int i, int) { return m.coeffRef(i, i); }
    using Base::coeff;
    using Base::coeffRef;
};

// ------------------------------------------------------------------------------------------------
// `m << a, b, c, ...`: scalars fill row-major; a column vector target also accepts sub-vectors, and a
// matrix target accepts blocks that are laid side by side row band by row band.
template <class XprType>
class CommaInitializer {
    XprType& x;
    int row = 0, col = 0, band = 1;

public:
    typedef typename traits<XprType>::Scalar Scalar;
    explicit CommaInitializer(XprType& xx) : x(xx) {}
    CommaInitializer& put(Scalar s) {
        if (col == x.cols()) {
            row += band;
            col = 0;
            band = 1;
        }
        assert(row < x.rows());
        x.coeffRef(row, col++) = s;
        return *this;
    }
    template <class O>
    CommaInitializer& put(const MatrixBase<O>& o) {
        if (col == x.cols()) {
            row += band;
            col = 0;
            band = 1;
        }
        if (col == 0) band = o.rows();
        assert(o.rows() == band && row + o.rows() <= x.rows() && col + o.cols() <= x.cols());
        for (int j = 0; j < o.cols(); ++j)
            for (int i = 0; i < o.rows(); ++i) x.coeffRef(row + i, col + j) = o.coeff(i, j);
        col += o.cols();
        return *this;
    }
    CommaInitializer& operator,(Scalar s) { return put(s); }
    template <class O>
    CommaInitializer& operator,(const MatrixBase<O>& o) {
        return put(o);
    }
    XprType& finished() { return x; }
};
template <class D>
CommaInitializer<D> MatrixBase<D>::operator<<(Scalar s) {
    CommaInitializer<D> ci(derived());
    ci.put(s);
    return ci;
}
template <class D>
template <class O>
CommaInitializer<D> MatrixBase<D>::operator<<(const MatrixBase<O>& o) {
    CommaInitializer<D> ci(derived());
    ci.put(o);
    return ci;
}

// ------------------------------------------------------------------------------------------------
// Arithmetic.  Every result is evaluated at once; inner products accumulate k = 0,1,2,... in order.
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> operator*(const MatrixBase<A>& a,
                                                                                const MatrixBase<B>& b) {
    typedef typename traits<A>::Scalar S;
    assert(a.cols() == b.rows());
    Matrix<S, traits<A>::Rows, traits<B>::Cols> out(a.rows(), b.cols());
    const int n = a.rows(), m = b.cols(), kk = a.cols();
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < n; ++i) {
            S s = S(0);
            for (int k = 0; k < kk; ++k) s += a.coeff(i, k) * b.coeff(k, j);
            out.coeffRef(i, j) = s;
        }
    return out;
}
template <class A, class B>
typename MatrixBase<A>::PlainObject operator+(const MatrixBase<A>& a, const MatrixBase<B>& b) {
    assert(a.rows() == b.rows() && a.cols() == b.cols());
    typename MatrixBase<A>::PlainObject out(a.rows(), a.cols());
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) out.coeffRef(i, j) = a.coeff(i, j) + b.coeff(i, j);
    return out;
}
template <class A, class B>
typename MatrixBase<A>::PlainObject operator-(const MatrixBase<A>& a, const MatrixBase<B>& b) {
    assert(a.rows() == b.rows() && a.cols() == b.cols());
    typename MatrixBase<A>::PlainObject out(a.rows(), a.cols());
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) out.coeffRef(i, j) = a.coeff(i, j) - b.coeff(i, j);
    return out;
}
template <class A>
typename MatrixBase<A>::PlainObject operator*(const MatrixBase<A>& a, typename traits<A>::Scalar s) {
    typename MatrixBase<A>::PlainObject out(a.rows(), a.cols());
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) out.coeffRef(i, j) = a.coeff(i, j) * s;
    return out;
}
template <class A>
typename MatrixBase<A>::PlainObject operator*(typename traits<A>::Scalar s, const MatrixBase<A>& a) {
    typename MatrixBase<A>::PlainObject out(a.rows(), a.cols());
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) out.coeffRef(i, j) = s * a.coeff(i, j);
    return out;
}
template <class A>
typename MatrixBase<A>::PlainObject operator/(const MatrixBase<A>& a, typename traits<A>::Scalar s) {
    typename MatrixBase<A>::PlainObject out(a.rows(), a.cols());
    for (int j = 0; j < a.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) out.coeffRef(i, j) = a.coeff(i, j) / s;
    return out;
}

template <class D>
std::ostream& operator<<(std::ostream& os, const MatrixBase<D>& m) {
    for (int i = 0; i < m.rows(); ++i) {
        for (int j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m.coeff(i, j);
        if (i + 1 < m.rows()) os << "\n";
    }
    return os;
}

// ------------------------------------------------------------------------------------------------
// In-place Gauss-Jordan solve with partial pivoting on column-major n x n `a` and n x m `b`.
template <class S>
inline bool mini_lu_solve(int n, int m, S* a, S* b) {
    for (int k = 0; k < n; ++k) {
        int p = k;
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(a[(size_t)k * n + i]) > std::fabs(a[(size_t)k * n + p])) p = i;
        if (a[(size_t)k * n + p] == S(0)) return false;
        if (p != k) {
            for (int j = 0; j < n; ++j) std::swap(a[(size_t)j * n + k], a[(size_t)j * n + p]);
            for (int j = 0; j < m; ++j) std::swap(b[(size_t)j * n + k], b[(size_t)j * n + p]);
        }
        for (int i = k + 1; i < n; ++i) {
            S f = a[(size_t)k * n + i] / a[(size_t)k * n + k];
            if (f == S(0)) continue;
            for (int j = k + 1; j < n; ++j) a[(size_t)j * n + i] -= f * a[(size_t)j * n + k];
            for (int j = 0; j < m; ++j) b[(size_t)j * n + i] -= f * b[(size_t)j * n + k];
        }
    }
    for (int j = 0; j < m; ++j)
        for (int i = n - 1; i >= 0; --i) {
            S s = b[(size_t)j * n + i];
            for (int k = i + 1; k < n; ++k) s -= a[(size_t)k * n + i] * b[(size_t)j * n + k];
            b[(size_t)j * n + i] = s / a[(size_t)i * n + i];
        }
    return true;
}

template <class D>
typename MatrixBase<D>::PlainObject MatrixBase<D>::inverse() const {
    assert(rows() == cols());
    PlainObject out(rows(), cols());
    if (rows() == 3) {
        // cofactors of the first column give the determinant; inverse = adjugate / det
        const MatrixBase& m = *this;
        auto cof = [&](int i, int j) {
            int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            return m.coeff(i1, j1) * m.coeff(i2, j2) - m.coeff(i1, j2) * m.coeff(i2, j1);
        };
        Scalar c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
        Scalar det = c0 * m.coeff(0, 0) + c1 * m.coeff(1, 0) + c2 * m.coeff(2, 0);
        Scalar invdet = Scalar(1) / det;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) out.coeffRef(j, i) = cof(i, j) * invdet;
        return out;
    }
    PlainObject a = eval();
    out.setIdentity();
    mini_lu_solve<Scalar>(rows(), rows(), a.data(), out.data());
    return out;
}

// `m.colwise() - v` / `m.colwise() + v` (v broadcast over the columns), `m.rowwise().sum()`, and
// `A.colPivHouseholderQr().solve(b)` (here: Gaussian elimination with partial pivoting).
template <class M>
class ColwiseProxy {
    const M& m;

public:
    typedef typename traits<M>::Scalar Scalar;
    typedef Matrix<Scalar, traits<M>::Rows, traits<M>::Cols> Plain;
    explicit ColwiseProxy(const M& mm) : m(mm) {}
    template <class O>
    Plain operator-(const MatrixBase<O>& v) const {
        Plain out(m.rows(), m.cols());
        for (int j = 0; j < m.cols(); ++j)
            for (int i = 0; i < m.rows(); ++i) out.coeffRef(i, j) = m.coeff(i, j) - v.coeff(i);
        return out;
    }
    template <class O>
    Plain operator+(const MatrixBase<O>& v) const {
        Plain out(m.rows(), m.cols());
        for (int j = 0; j < m.cols(); ++j)
            for (int i = 0; i < m.rows(); ++i) out.coeffRef(i, j) = m.coeff(i, j) + v.coeff(i);
        return out;
    }
};
template <class M>
class RowwiseProxy {
    const M& m;

public:
    typedef typename traits<M>::Scalar Scalar;
    explicit RowwiseProxy(const M& mm) : m(mm) {}
    Matrix<Scalar, traits<M>::Rows, 1> sum() const {
        Matrix<Scalar, traits<M>::Rows, 1> out(m.rows(), 1);
        for (int i = 0; i < m.rows(); ++i) {
            Scalar s = 0;
            for (int j = 0; j < m.cols(); ++j) s += m.coeff(i, j);
            out.coeffRef(i, 0) = s;
        }
        return out;
    }
};
template <class M>
class QrSolveProxy {
    const M& m;

public:
    typedef typename traits<M>::Scalar Scalar;
    explicit QrSolveProxy(const M& mm) : m(mm) {}
    template <class O>
    typename MatrixBase<O>::PlainObject solve(const MatrixBase<O>& b) const {
        Matrix<Scalar, Dynamic, Dynamic> a = m.eval();
        typename MatrixBase<O>::PlainObject x = b.eval();
        mini_lu_solve<Scalar>(a.rows(), x.cols(), a.data(), x.data());
        return x;
    }
};

// ------------------------------------------------------------------------------------------------
// Geometry: just the quaternion members the MPC path touches.
template <class T>
class Quaternion {
    T c[4];   // x y z w

public:
    typedef T Scalar;
    Quaternion() { c[0] = c[1] = c[2] = 0, c[3] = 1; }
    Quaternion(T w, T x, T y, T z) { c[0] = x, c[1] = y, c[2] = z, c[3] = w; }
    T& x() { return c[0]; }
    T& y() { return c[1]; }
    T& z() { return c[2]; }
    T& w() { return c[3]; }
    T x() const { return c[0]; }
    T y() const { return c[1]; }
    T z() const { return c[2]; }
    T w() const { return c[3]; }
    T squaredNorm() const { return c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3]; }
    Quaternion inverse() const {
        T n2 = squaredNorm();
        return Quaternion(c[3] / n2, -c[0] / n2, -c[1] / n2, -c[2] / n2);
    }
    Matrix<T, 3, 3> toRotationMatrix() const {
        Matrix<T, 3, 3> res;
        const T tx = T(2) * x(), ty = T(2) * y(), tz = T(2) * z();
        const T twx = tx * w(), twy = ty * w(), twz = tz * w();
        const T txx = tx * x(), txy = ty * x(), txz = tz * x();
        const T tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
        res.coeffRef(0, 0) = T(1) - (tyy + tzz);
        res.coeffRef(0, 1) = txy - twz;
        res.coeffRef(0, 2) = txz + twy;
        res.coeffRef(1, 0) = txy + twz;
        res.coeffRef(1, 1) = T(1) - (txx + tzz);
        res.coeffRef(1, 2) = tyz - twx;
        res.coeffRef(2, 0) = txz - twy;
        res.coeffRef(2, 1) = tyz + twx;
        res.coeffRef(2, 2) = T(1) - (txx + tyy);
        return res;
    }
};
typedef Quaternion<float> Quaternionf;
typedef Quaternion<double> Quaterniond;

// Thin singular value decomposition by one-sided (Hestenes) Jacobi rotations on the columns of the
// taller of (A, A^T); singular values sorted in decreasing order.  Eigen's JacobiSVD is two-sided with
// a QR preconditioner: same factorisation, different rounding.
template <class M>
class JacobiSVD {
public:
    typedef typename M::Scalar Scalar;
    typedef Matrix<Scalar, Dynamic, 1> SingularValuesType;
    JacobiSVD() {}
    JacobiSVD(const M& a, unsigned = 0) { compute(a); }
    JacobiSVD& compute(const M& a, unsigned = 0) {
        const bool wide = a.rows() < a.cols();
        M B = wide ? M(a.transpose()) : a;   // n x k, n >= k
        const int n = B.rows(), k = B.cols();
        M V(k, k);
        V.setIdentity();
        const double eps = sizeof(Scalar) == 4 ? 1e-7 : 1e-15;
        for (int sweep = 0; sweep < 60; ++sweep) {
            double off = 0;
            for (int p = 0; p < k - 1; ++p)
                for (int q = p + 1; q < k; ++q) {
                    double alpha = 0, beta = 0, gamma = 0;
                    for (int i = 0; i < n; ++i) {
                        alpha += double(B.coeff(i, p)) * B.coeff(i, p);
                        beta += double(B.coeff(i, q)) * B.coeff(i, q);
                        gamma += double(B.coeff(i, p)) * B.coeff(i, q);
                    }
                    if (gamma == 0) continue;
                    double rel = std::fabs(gamma) / std::sqrt(alpha * beta + 1e-300);
                    if (rel > off) off = rel;
                    double zeta = (beta - alpha) / (2 * gamma);
                    double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
                    double c = 1 / std::sqrt(1 + t * t), sn = c * t;
                    for (int i = 0; i < n; ++i) {
                        Scalar bp = B.coeff(i, p), bq = B.coeff(i, q);
                        B.coeffRef(i, p) = Scalar(c * bp - sn * bq);
                        B.coeffRef(i, q) = Scalar(sn * bp + c * bq);
                    }
                    for (int i = 0; i < k; ++i) {
                        Scalar vp = V.coeff(i, p), vq = V.coeff(i, q);
                        V.coeffRef(i, p) = Scalar(c * vp - sn * vq);
                        V.coeffRef(i, q) = Scalar(sn * vp + c * vq);
                    }
                }
            if (off < eps) break;
        }
        std::vector<double> sig(k);
        std::vector<int> order(k);
        for (int j = 0; j < k; ++j) {
            double s2 = 0;
            for (int i = 0; i < n; ++i) s2 += double(B.coeff(i, j)) * B.coeff(i, j);
            sig[j] = std::sqrt(s2);
            order[j] = j;
        }
        for (int a1 = 0; a1 < k; ++a1)   // selection sort, stable enough for k <= 18
            for (int b1 = a1 + 1; b1 < k; ++b1)
                if (sig[order[b1]] > sig[order[a1]]) std::swap(order[a1], order[b1]);
        M Ub(n, k), Vs(k, k);
        s_.resize(k);
        for (int j = 0; j < k; ++j) {
            const int o = order[j];
            s_.coeffRef(j) = Scalar(sig[o]);
            for (int i = 0; i < n; ++i) Ub.coeffRef(i, j) = sig[o] > 0 ? Scalar(B.coeff(i, o) / sig[o]) : Scalar(0);
            for (int i = 0; i < k; ++i) Vs.coeffRef(i, j) = V.coeff(i, o);
        }
        if (wide) {   // A^T = Ub S Vs^T  =>  A = Vs S Ub^T
            u_ = Vs;
            v_ = Ub;
        } else {
            u_ = Ub;
            v_ = Vs;
        }
        return *this;
    }
    const SingularValuesType& singularValues() const { return s_; }
    const M& matrixU() const { return u_; }
    const M& matrixV() const { return v_; }

private:
    M u_, v_;
    SingularValuesType s_;
};

// Linear solves the reference spells as a rank-revealing QR; here: LU with partial pivoting.
template <class M>
class ColPivHouseholderQR {
    M a;

public:
    ColPivHouseholderQR() {}
    explicit ColPivHouseholderQR(const M& m) : a(m) {}
    ColPivHouseholderQR& compute(const M& m) {
        a = m;
        return *this;
    }
    template <class B>
    typename MatrixBase<B>::PlainObject solve(const MatrixBase<B>& b) const {
        M lu = a;
        typename MatrixBase<B>::PlainObject x = b.eval();
        mini_lu_solve<typename M::Scalar>(lu.rows(), x.cols(), lu.data(), x.data());
        return x;
    }
};

template <class T, int Dim, int Mode>
class Transform {
    // Isometry3 with Eigen's semantics: translate() and rotate() post-multiply (T <- T * Translation(v), T <- T * R),
    // a point is mapped to linear * p + translation.
    Matrix<T, 3, 3> lin;
    Matrix<T, 3, 1> tr;

public:
    Transform() {
        lin.setIdentity();
        tr.setZero();
    }
    static Transform Identity() { return Transform(); }
    template <class V>
    Transform& translate(const V& v) {
        Matrix<T, 3, 1> vv;
        for (int i = 0; i < 3; ++i) vv(i) = v[i];
        Matrix<T, 3, 1> d = lin * vv;
        tr += d;
        return *this;
    }
    Transform& rotate(const Quaternion<T>& q) {
        Matrix<T, 3, 3> r = q.toRotationMatrix();
        lin = (lin * r).eval();
        return *this;
    }
    const Matrix<T, 3, 3>& linear() const { return lin; }
    const Matrix<T, 3, 1>& translation() const { return tr; }
    template <class P>
    P operator*(const P& p) const {
        P out = p;
        for (int j = 0; j < p.cols(); ++j)
            for (int i = 0; i < 3; ++i) {
                T s = 0;
                for (int k = 0; k < 3; ++k) s += lin(i, k) * p(k, j);
                out(i, j) = s + tr(i);
            }
        return out;
    }
};

}   // namespace Eigen

#include "mini_eigen_expm.h"
#endif
