// placeholder, filled in below
#include "orb_oracle.h"
