// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for gtl v1.2.0 (pinned by the reference at
// cmake/gtl.cmake:8; used at include/chad/detail/octree.hpp:187 and levels.hpp:18,93,143).
// The reference never iterates a gtl container, so results do not depend on the hash or the
// probing scheme; only speed does. Both containers here are linear-probing open-addressing
// tables (load <= 0.5, 64-bit mix of the user hash) so the CPU baseline is not handicapped
// by node-based std::unordered_*.
// The reference relies on gtl for these transitive includes (levels.hpp:14,39,57).
#pragma once
#include <array>
#include <bit>
#include <cstdint>
#include <cstring>
#include <functional>
#include <utility>
#include <vector>

namespace gtl {
namespace shim_detail {
inline uint64_t mix(uint64_t h) {
    h ^= h >> 32;
    h *= 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    return h;
}
}  // namespace shim_detail

struct HashState {
    template <typename T>
    static uint64_t combine(uint64_t seed, const T& v) {
        return seed ^ (shim_detail::mix(uint64_t(v)) + 0x9E3779B97F4A7C15ull + (seed << 6) + (seed >> 2));
    }
};

template <typename K, typename V, typename Hash = std::hash<K>>
class flat_hash_map {
    struct Slot { K first; V second; };
    struct Cell { bool used; alignas(Slot) unsigned char raw[sizeof(Slot)]; };

public:
    using value_type = Slot;
    struct iterator {
        Slot* p;
        Slot* operator->() const { return p; }
        Slot& operator*() const { return *p; }
    };
    flat_hash_map() { rebuild(1u << 12); }
    template <typename KK, typename VV>
    std::pair<iterator, bool> try_emplace(const KK& key_in, VV&& val) {
        if ((_size + 1) * 2 > _cells.size()) grow();
        K key(key_in);
        size_t i = shim_detail::mix(_hash(key)) & _mask;
        while (true) {
            Cell& c = _cells[i];
            if (!c.used) {
                c.used = true;
                Slot* s = new (c.raw) Slot{key, V(std::forward<VV>(val))};
                _size++;
                return {iterator{s}, true};
            }
            Slot* s = reinterpret_cast<Slot*>(c.raw);
            if (s->first == key) return {iterator{s}, false};
            i = (i + 1) & _mask;
        }
    }
    void clear() {
        for (auto& c : _cells) c.used = false;
        _size = 0;
    }
    size_t size() const { return _size; }

private:
    void rebuild(size_t n) {
        _cells.assign(n, Cell{});
        _mask = n - 1;
        _size = 0;
    }
    void grow() {
        std::vector<Cell> old;
        old.swap(_cells);
        rebuild(old.size() * 2);
        for (auto& c : old) {
            if (!c.used) continue;
            Slot* s = reinterpret_cast<Slot*>(c.raw);
            size_t i = shim_detail::mix(_hash(s->first)) & _mask;
            while (_cells[i].used) i = (i + 1) & _mask;
            _cells[i].used = true;
            new (_cells[i].raw) Slot{s->first, s->second};
            _size++;
        }
    }
    std::vector<Cell> _cells;
    size_t _mask = 0, _size = 0;
    Hash _hash;
};

// As used by the reference (default null mutex) this is a plain, single-threaded set.
template <typename K, typename Hash, typename Eq>
class parallel_flat_hash_set {
    static constexpr K kEmpty = K(~K(0));

public:
    struct iterator {
        const K* p;
        const K& operator*() const { return *p; }
    };
    parallel_flat_hash_set(size_t /*bucket_count*/, const Hash& h, const Eq& e) : _hash(h), _eq(e) { rebuild(1u << 12); }
    std::pair<iterator, bool> emplace(const K& key) {
        if ((_size + 1) * 2 > _keys.size()) grow();
        size_t i = shim_detail::mix(_hash(key)) & _mask;
        while (true) {
            if (_keys[i] == kEmpty) {
                _keys[i] = key;
                _size++;
                return {iterator{&_keys[i]}, true};
            }
            if (_eq(_keys[i], key)) return {iterator{&_keys[i]}, false};
            i = (i + 1) & _mask;
        }
    }
    size_t size() const { return _size; }

private:
    void rebuild(size_t n) {
        _keys.assign(n, kEmpty);
        _mask = n - 1;
        _size = 0;
    }
    void grow() {
        std::vector<K> old;
        old.swap(_keys);
        rebuild(old.size() * 2);
        for (K k : old) {
            if (k == kEmpty) continue;
            size_t i = shim_detail::mix(_hash(k)) & _mask;
            while (_keys[i] != kEmpty) i = (i + 1) & _mask;
            _keys[i] = k;
            _size++;
        }
    }
    std::vector<K> _keys;
    size_t _mask = 0, _size = 0;
    Hash _hash;
    Eq _eq;
};
}  // namespace gtl
