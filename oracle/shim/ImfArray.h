// oracle/shim/ImfArray.h — TEST INFRASTRUCTURE ONLY. Imf::Array2D as used by
// reference src/bitmap.cpp:246-250 (resizeErase + operator[] row access).
#pragma once
#include <vector>
namespace Imf {
template <class T> class Array2D {
    std::vector<T> m_data;
    long m_sx = 0, m_sy = 0;
public:
    void resizeErase(long sizeX, long sizeY) { m_sx = sizeX; m_sy = sizeY; m_data.assign((size_t)sizeX * sizeY, T()); }
    T* operator[](long x) { return m_data.data() + x * m_sy; }
    const T* operator[](long x) const { return m_data.data() + x * m_sy; }
};
}
