// mfem.hpp -- MINIMAL STAND-IN for the MFEM 4.7 dense-algebra / element API, written
// from the documented behaviour of those calls.  TEST INFRASTRUCTURE ONLY.
//
// Purpose: MFEM itself is absent from this image, but the arithmetic of the hot path
// lives in the reference's OWN file MFEM/mechanic2d/asym_elasto_damage_model.cc
// (asym_stress, lines 207-329; damIntegrator, lines 490-953).  build_ref.sh compiles
// exactly those line ranges, from where they lie under /root/reference, against this
// header, so that the oracle (oracle/fem_oracle.c) can be checked against the
// reference's own code and golden vectors can be generated from it
// (tests/golden/make_ref_vectors.py).  Only what those line ranges use is provided.
//
// Semantics implemented (MFEM 4.7): DenseMatrix is column-major; Mult(A,B,C): C = A B;
// MultAtB: A^t B; MultADAt: A diag(D) A^t; AddMult: C += A B; AddMult_a_VWt: M += a v w^t;
// AddMult_a_VVt: M += a v v^t; AddMult_a_ABt: M += a A B^t; Add(A,B,alpha,C): C = A +
// alpha B; Add(alpha,A,beta,B,C): C = alpha A + beta B; Linear2DFiniteElement::CalcDShape
// = [[-1,-1],[1,0],[0,1]], CalcShape = [1-x-y, x, y]; IsoparametricTransformation:
// J = PointMat * dshape, Weight() = det J (signed), InverseJacobian() = J^-1.
#pragma once
#include <cassert>
#include <chrono>
#include <cmath>
#include <iostream>
#include <vector>

namespace mfem {

typedef double real_t;

// MFEM_ASSERT is compiled out unless MFEM_DEBUG is defined; the reference builds with
// -O3 -DNDEBUG against a release MFEM (MFEM/setting.mk.in:4), so it is a no-op here too
#define MFEM_ASSERT(cond, msg)
#define MFEM_VERIFY(cond, msg) \
   if (!(cond)) { std::cerr << "MFEM_VERIFY failed: " << #cond << std::endl; std::abort(); }

class Vector
{
   double *data;
   int size;
   std::vector<double> own;

  public:
   Vector() : data(nullptr), size(0) {}
   explicit Vector(int n) : size(n), own(n, 0.) { data = own.data(); }
   Vector(double *d, int n) : data(d), size(n) {}
   Vector(const Vector &o) : size(o.size), own(o.data, o.data + o.size) { data = own.data(); }
   Vector &operator=(const Vector &o)
   {  // deep copy of the values (sizes follow the source)
      SetSize(o.size);
      for (int i = 0; i < size; ++i) data[i] = o.data[i];
      return *this;
   }
   void SetSize(int n)
   {
      if (n == size) return;
      own.assign(n, 0.);
      data = own.data();
      size = n;
   }
   int Size() const { return size; }
   double *GetData() const { return data; }
   double &operator[](int i) { return data[i]; }
   const double &operator[](int i) const { return data[i]; }
   double &operator()(int i) { return data[i]; }
   const double &operator()(int i) const { return data[i]; }
   Vector &operator=(double v)
   {
      for (int i = 0; i < size; ++i) data[i] = v;
      return *this;
   }
   void Print(std::ostream &os = std::cout) const
   {
      for (int i = 0; i < size; ++i) os << data[i] << ' ';
      os << '\n';
   }
};

class DenseMatrix
{
   double *data;
   int h, w;
   std::vector<double> own;

  public:
   DenseMatrix() : data(nullptr), h(0), w(0) {}
   explicit DenseMatrix(int n) : h(n), w(n), own((size_t)n * n, 0.) { data = own.data(); }
   DenseMatrix(int m, int n) : h(m), w(n), own((size_t)m * n, 0.) { data = own.data(); }
   DenseMatrix(const DenseMatrix &o) : h(o.h), w(o.w), own(o.data, o.data + (size_t)o.h * o.w) { data = own.data(); }
   DenseMatrix &operator=(const DenseMatrix &o)
   {  // deep copy of the values (sizes follow the source)
      SetSize(o.h, o.w);
      for (int i = 0; i < h * w; ++i) data[i] = o.data[i];
      return *this;
   }
   void SetSize(int n) { SetSize(n, n); }
   void SetSize(int m, int n)
   {
      if (m == h && n == w) return;
      own.assign((size_t)m * n, 0.);
      data = own.data();
      h = m, w = n;
   }
   void UseExternalData(double *d, int m, int n) { data = d, h = m, w = n; }
   int Height() const { return h; }
   int Width() const { return w; }
   double *GetData() const { return data; }
   double &operator()(int i, int j) { return data[i + (size_t)j * h]; }
   const double &operator()(int i, int j) const { return data[i + (size_t)j * h]; }
   DenseMatrix &operator=(double v)
   {
      for (int i = 0; i < h * w; ++i) data[i] = v;
      return *this;
   }
   DenseMatrix &operator*=(double c)
   {
      for (int i = 0; i < h * w; ++i) data[i] *= c;
      return *this;
   }
   void Symmetrize()
   {
      for (int i = 0; i < h; ++i)
         for (int j = 0; j < i; ++j)
         {
            const double a = 0.5 * ((*this)(i, j) + (*this)(j, i));
            (*this)(i, j) = (*this)(j, i) = a;
         }
   }
   void Norm2(double *v) const
   {  // Euclidean norm of every column
      for (int j = 0; j < w; ++j)
      {
         double s = 0.;
         for (int i = 0; i < h; ++i) s += (*this)(i, j) * (*this)(i, j);
         v[j] = std::sqrt(s);
      }
   }
   void SetSubMatrix(int ro, int co, const DenseMatrix &A)
   {
      for (int j = 0; j < A.w; ++j)
         for (int i = 0; i < A.h; ++i) (*this)(ro + i, co + j) = A(i, j);
   }
   void AddSubMatrix(int ro, int co, const DenseMatrix &A)
   {
      for (int j = 0; j < A.w; ++j)
         for (int i = 0; i < A.h; ++i) (*this)(ro + i, co + j) += A(i, j);
   }
   void AddSubMatrix(int o, const DenseMatrix &A) { AddSubMatrix(o, o, A); }
   void Transpose()
   {
      assert(h == w);
      for (int i = 0; i < h; ++i)
         for (int j = 0; j < i; ++j) std::swap((*this)(i, j), (*this)(j, i));
   }
   void Transpose(const DenseMatrix &A)
   {
      SetSize(A.w, A.h);
      for (int i = 0; i < A.h; ++i)
         for (int j = 0; j < A.w; ++j) (*this)(j, i) = A(i, j);
   }
   DenseMatrix &Add(const double c, const DenseMatrix &A)
   {
      for (int i = 0; i < h * w; ++i) data[i] += c * A.data[i];
      return *this;
   }
   void Print(std::ostream &os = std::cout) const
   {
      for (int i = 0; i < h; ++i)
      {
         for (int j = 0; j < w; ++j) os << (*this)(i, j) << ' ';
         os << '\n';
      }
   }
};

inline void Mult(const DenseMatrix &A, const DenseMatrix &B, DenseMatrix &C)
{
   assert(A.Width() == B.Height() && C.Height() == A.Height() && C.Width() == B.Width());
   for (int j = 0; j < B.Width(); ++j)
      for (int i = 0; i < A.Height(); ++i)
      {
         double s = 0.;
         for (int k = 0; k < A.Width(); ++k) s += A(i, k) * B(k, j);
         C(i, j) = s;
      }
}
inline void AddMult(const DenseMatrix &A, const DenseMatrix &B, DenseMatrix &C)
{
   for (int j = 0; j < B.Width(); ++j)
      for (int i = 0; i < A.Height(); ++i)
      {
         double s = 0.;
         for (int k = 0; k < A.Width(); ++k) s += A(i, k) * B(k, j);
         C(i, j) += s;
      }
}
inline void MultAtB(const DenseMatrix &A, const DenseMatrix &B, DenseMatrix &AtB)
{
   assert(A.Height() == B.Height() && AtB.Height() == A.Width() && AtB.Width() == B.Width());
   for (int j = 0; j < B.Width(); ++j)
      for (int i = 0; i < A.Width(); ++i)
      {
         double s = 0.;
         for (int k = 0; k < A.Height(); ++k) s += A(k, i) * B(k, j);
         AtB(i, j) = s;
      }
}
inline void MultADAt(const DenseMatrix &A, const Vector &D, DenseMatrix &ADAt)
{
   for (int i = 0; i < A.Height(); ++i)
      for (int j = 0; j < A.Height(); ++j)
      {
         double s = 0.;
         for (int k = 0; k < A.Width(); ++k) s += A(i, k) * D[k] * A(j, k);
         ADAt(i, j) = s;
      }
}
inline void AddMult_a_VWt(const double a, const Vector &v, const Vector &w, DenseMatrix &VWt)
{
   for (int j = 0; j < w.Size(); ++j)
      for (int i = 0; i < v.Size(); ++i) VWt(i, j) += a * v[i] * w[j];
}
inline void AddMult_a_VVt(const double a, const Vector &v, DenseMatrix &VVt)
{
   for (int j = 0; j < v.Size(); ++j)
      for (int i = 0; i < v.Size(); ++i) VVt(i, j) += a * v[i] * v[j];
}
inline void AddMult_a_ABt(double a, const DenseMatrix &A, const DenseMatrix &B, DenseMatrix &ABt)
{
   for (int j = 0; j < B.Height(); ++j)
      for (int i = 0; i < A.Height(); ++i)
      {
         double s = 0.;
         for (int k = 0; k < A.Width(); ++k) s += A(i, k) * B(j, k);
         ABt(i, j) += a * s;
      }
}
inline void Add(const DenseMatrix &A, const DenseMatrix &B, double alpha, DenseMatrix &C)
{
   for (int j = 0; j < C.Width(); ++j)
      for (int i = 0; i < C.Height(); ++i) C(i, j) = A(i, j) + alpha * B(i, j);
}
inline void Add(double alpha, const DenseMatrix &A, double beta, const DenseMatrix &B, DenseMatrix &C)
{
   for (int j = 0; j < C.Width(); ++j)
      for (int i = 0; i < C.Height(); ++i) C(i, j) = alpha * A(i, j) + beta * B(i, j);
}

struct IntegrationPoint
{
   double x = 0., y = 0., z = 0., weight = 0.;
   int index = 0;
};
class IntegrationRule
{
   std::vector<IntegrationPoint> pts;

  public:
   IntegrationRule() {}
   explicit IntegrationRule(int n) : pts(n) {}
   int GetNPoints() const { return (int)pts.size(); }
   IntegrationPoint &IntPoint(int i) { return pts[i]; }
   const IntegrationPoint &IntPoint(int i) const { return pts[i]; }
};

// straight-sided triangle, P1 geometry (IsoparametricTransformation of a Linear2DFiniteElement)
class ElementTransformation
{
   double X[3][2];
   DenseMatrix invJ;
   double det;
   const IntegrationPoint *ip;

  public:
   int Attribute = 1, ElementNo = 0;
   explicit ElementTransformation(const double *xv) : invJ(2), ip(nullptr)
   {
      for (int a = 0; a < 3; ++a) X[a][0] = xv[2 * a], X[a][1] = xv[2 * a + 1];
      // J(i, j) = sum_a X[a][i] dshape(a, j), dshape = [[-1,-1],[1,0],[0,1]]
      const double J00 = X[1][0] - X[0][0], J01 = X[2][0] - X[0][0];
      const double J10 = X[1][1] - X[0][1], J11 = X[2][1] - X[0][1];
      det = J00 * J11 - J01 * J10;
      invJ(0, 0) = J11 / det, invJ(0, 1) = -J01 / det;
      invJ(1, 0) = -J10 / det, invJ(1, 1) = J00 / det;
   }
   void SetIntPoint(const IntegrationPoint *p) { ip = p; }
   const IntegrationPoint &GetIntPoint() const { return *ip; }
   double Weight() const { return det; }
   const DenseMatrix &InverseJacobian() const { return invJ; }
   int GetSpaceDim() const { return 2; }
};

class FiniteElement
{
  public:
   int GetDof() const { return 3; }
   int GetDim() const { return 2; }
   void CalcDShape(const IntegrationPoint &, DenseMatrix &dshape) const
   {
      dshape(0, 0) = -1., dshape(0, 1) = -1.;
      dshape(1, 0) = 1., dshape(1, 1) = 0.;
      dshape(2, 0) = 0., dshape(2, 1) = 1.;
   }
   void CalcPhysShape(ElementTransformation &Tr, Vector &shape) const
   {
      const IntegrationPoint &p = Tr.GetIntPoint();
      shape[0] = 1. - p.x - p.y, shape[1] = p.x, shape[2] = p.y;
   }
};

class Coefficient
{
  public:
   virtual double Eval(ElementTransformation &T, const IntegrationPoint &ip) = 0;
   virtual ~Coefficient() {}
};
class ConstantCoefficient : public Coefficient
{
  public:
   double constant;
   explicit ConstantCoefficient(double c = 1.) : constant(c) {}
   double Eval(ElementTransformation &, const IntegrationPoint &) override { return constant; }
};
// value at the (single) quadrature point of the element
class QuadratureFunctionCoefficient : public ConstantCoefficient
{
  public:
   explicit QuadratureFunctionCoefficient(double v = 0.) : ConstantCoefficient(v) {}
};
// values at the quadrature points of the element, indexed by IntegrationPoint::index
class VectorQuadratureFunctionCoefficient
{
  public:
   std::vector<double> values;  // [npoints][2]
   void Eval(Vector &V, ElementTransformation &, const IntegrationPoint &ip)
   {
      V[0] = values[2 * ip.index], V[1] = values[2 * ip.index + 1];
   }
};

class NonlinearFormIntegrator
{
  protected:
   const IntegrationRule *IntRule;

  public:
   explicit NonlinearFormIntegrator(const IntegrationRule *ir = nullptr) : IntRule(ir) {}
   virtual ~NonlinearFormIntegrator() {}
};

}  // namespace mfem

inline double MPI_Wtime()
{
   return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
