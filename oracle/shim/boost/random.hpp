// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the subset of Boost.Random 1.64.0 that
// the reference driver uses (reference code/jam/jamming.cpp:36-41, :882-887). Boost is not
// vendored by the reference and not installed here, so only the RNG plumbing is substituted;
// every line of physics still comes from the reference's own sources.
//
// What is mirrored (from the published Boost 1.64 algorithm; parity of RNG *streams* with a
// real Boost build is unpinned -- the reference seeds from the wall clock, jamming.cpp:36-37,
// so no reference run is reproducible anyway):
//   * boost::mt19937            -> std::mt19937 (same MT19937 recurrence, 32-bit output)
//   * boost::uniform_real<>     -> u32 / 2^32 * (max - min) + min   (generate_uniform_real)
//   * boost::normal_distribution<> -> std::normal_distribution (Boost uses a ziggurat; only
//                                     the distribution matters for radii / lattice jitter)
//   * boost::variate_generator<E, D>  by-value engine copy (jamming.cpp:40-41),
//                                     and the E& specialisation (jamming.cpp:885).
//
// Added for the parity gates: a variate_generator can be put in "injection" mode, in which
// operator() returns caller-supplied values in call order (the reference draws exactly one
// randuni() per particle per step in ascending particle index, jamming.cpp:667).
#pragma once
#include <cstddef>
#include <cstdint>
#include <random>

namespace boost {

class mt19937 {
public:
    typedef std::uint32_t result_type;
    mt19937() : eng_(5489u) {}
    template <class T> explicit mt19937(T seed) : eng_(static_cast<std::uint32_t>(seed)) {}
    result_type operator()() { return static_cast<result_type>(eng_()); }
    void seed(std::uint32_t s) { eng_.seed(s); }
    static constexpr result_type min() { return 0u; }
    static constexpr result_type max() { return 0xffffffffu; }
private:
    std::mt19937 eng_;
};

template <class RealType = double>
class uniform_real {
public:
    typedef RealType result_type;
    explicit uniform_real(RealType lo = RealType(0), RealType hi = RealType(1)) : lo_(lo), hi_(hi) {}
    template <class Engine> result_type operator()(Engine& eng) {
        for (;;) {
            RealType numerator = static_cast<RealType>(eng() - (eng.min)());
            RealType divisor = static_cast<RealType>((eng.max)() - (eng.min)()) + 1;
            RealType result = numerator / divisor * (hi_ - lo_) + lo_;
            if (result < hi_) return result;
        }
    }
    void reset() {}
private:
    RealType lo_, hi_;
};

template <class RealType = double>
class normal_distribution {
public:
    typedef RealType result_type;
    explicit normal_distribution(RealType mean = RealType(0), RealType sigma = RealType(1))
        : d_(mean, sigma) {}
    template <class Engine> result_type operator()(Engine& eng) { return d_(eng); }
    void reset() { d_.reset(); }
private:
    std::normal_distribution<RealType> d_;
};

namespace detail {
template <class E> struct engine_holder { E e; explicit engine_holder(const E& x) : e(x) {} E& get() { return e; } };
template <class E> struct engine_holder<E&> { E* e; explicit engine_holder(E& x) : e(&x) {} E& get() { return *e; } };
}  // namespace detail

template <class Engine, class Distribution>
class variate_generator {
public:
    typedef typename Distribution::result_type result_type;
    template <class E> variate_generator(E& e, Distribution d) : eng_(e), dist_(d) {}
    result_type operator()() {
        if (inject_ != nullptr && inject_pos_ < inject_n_) return inject_[inject_pos_++];
        return dist_(eng_.get());
    }
    // --- oracle-only hooks -------------------------------------------------------------
    void oracle_inject(const result_type* values, std::size_t n) { inject_ = values; inject_n_ = n; inject_pos_ = 0; }
    void oracle_clear_injection() { inject_ = nullptr; inject_n_ = inject_pos_ = 0; }
    std::size_t oracle_injected_consumed() const { return inject_pos_; }
    void oracle_reseed(std::uint32_t s) { eng_.get().seed(s); dist_.reset(); }
private:
    detail::engine_holder<Engine> eng_;
    Distribution dist_;
    const result_type* inject_ = nullptr;
    std::size_t inject_n_ = 0, inject_pos_ = 0;
};

}  // namespace boost
