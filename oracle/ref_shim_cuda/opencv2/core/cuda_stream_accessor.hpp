// stand-in: see cvcuda_shim.hpp
#pragma once
#include <cvcuda_shim.hpp>
