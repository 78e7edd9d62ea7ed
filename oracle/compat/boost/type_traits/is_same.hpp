#include "../type_traits.hpp"
