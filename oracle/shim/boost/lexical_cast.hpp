// TEST INFRASTRUCTURE ONLY (oracle/): the reference includes <boost/lexical_cast.hpp>
// (code/jam/jamming.cpp:23) but never uses it; an empty header satisfies the include.
#pragma once
