// Test driver for lethe_b200/host/function_expression.h: eval_expression x y z t 'expr' ['expr' ...]
// prints one value per expression with 17 significant digits, or "error: ..." for a parse error.
#include <cstdio>
#include <cstdlib>

#include "../../lethe_b200/host/function_expression.h"

int main(int argc, char **argv)
{
  if (argc < 6)
    return 2;
  lethe_b200::FunctionExpression::Variables v;
  v.x = std::atof(argv[1]);
  v.y = std::atof(argv[2]);
  v.z = std::atof(argv[3]);
  v.t = std::atof(argv[4]);
  for (int k = 5; k < argc; ++k)
    {
      try
        {
          const lethe_b200::FunctionExpression f(argv[k]);
          std::printf("%.17g %d\n", f(v), int(f.is_constant()));
        }
      catch (const std::exception &e)
        {
          std::printf("error: %s\n", e.what());
        }
    }
  return 0;
}
