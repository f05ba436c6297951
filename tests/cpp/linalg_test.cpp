// Dense helpers behind the mirror's approximateCovarianceBySampling (no GPU work).
// usage: linalg_test <in.bin> <out.bin>;  in: int32 m, n; A (m*n f64, row-major); c (m f64); H (9 f64)
//                                         out: int32 ok; q (n f64); eigenvalues of H (3 f64); inverse of H (9 f64)
#include <cstdio>
#include <vector>
#include "cfear_b200.hpp"
using namespace CFEAR_Radarodometry;
int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  int32_t mn[2];
  if (fread(mn, 4, 2, f) != 2) return 4;
  std::vector<double> A((size_t)mn[0] * mn[1]), c(mn[0]), q(mn[1], 0.0);
  double H[9], ev[3], Hi[9];
  if (fread(A.data(), 8, A.size(), f) != A.size() || fread(c.data(), 8, c.size(), f) != c.size() || fread(H, 8, 9, f) != 9) return 4;
  fclose(f);
  const int32_t ok = detail::lstsq_qr(A, mn[0], mn[1], c, q.data()) ? 1 : 0;
  detail::eig3_sym(H, ev);
  const int32_t ok2 = detail::inv3(H, Hi) ? 1 : 0;
  FILE* o = fopen(argv[2], "wb");
  const int32_t okk = ok & ok2;
  fwrite(&okk, 4, 1, o); fwrite(q.data(), 8, q.size(), o); fwrite(ev, 8, 3, o); fwrite(Hi, 8, 9, o);
  fclose(o);
  return 0;
}
