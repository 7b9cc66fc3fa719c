// stand-in: see g2o/stba_g2o.h (the g2o-shaped front door over the stba C ABI; NOT g2o)
#include "../stba_g2o.h"
