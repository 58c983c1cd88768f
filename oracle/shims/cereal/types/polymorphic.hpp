#include "../qadc_cereal_stub.hpp"
