// cf_matrix.h -- dense row-major table of T.
//
// INTERFACE-MANDATED by the callers written against the reference's matrix<T> (matrix.h:9-90): the name, rows() / cols(),
// m[i][j] through a row pointer, resize(rows, cols), begin() / end() over the rows * cols live cells, empty(), swap,
// construction and assignment from a matrix of another value type, and the free function transpose().  One behaviour is
// kept on purpose: resize never shrinks the storage and does not preserve the cells (callers fill after resizing).
// OWN STRUCTURE: cells are addressed through one private cell() helper; conversions go through a generic copy.
#pragma once

#include <algorithm>
#include <cstddef>
#include <vector>

template <class T>
class matrix
{
public:
    using iterator = typename std::vector<T>::iterator;
    using const_iterator = typename std::vector<T>::const_iterator;

    matrix() = default;
    matrix(const size_t rows, const size_t cols) : nRow(rows), nCol(cols), cells(rows * cols) {}

    // from a table of another value type (double <-> Number)
    template <class U>
    matrix(const matrix<U>& other) : matrix(other.rows(), other.cols())
    {
        std::transform(other.begin(), other.end(), cells.begin(), [](const U& x) { return T(x); });
    }
    template <class U>
    matrix& operator=(const matrix<U>& other)
    {
        matrix converted(other);
        swap(converted);
        return *this;
    }

    size_t rows() const { return nRow; }
    size_t cols() const { return nCol; }
    bool   empty() const { return cells.empty(); }

    T*       operator[](const size_t row) { return cell(row, 0); }
    const T* operator[](const size_t row) const { return cell(row, 0); }

    iterator       begin() { return cells.begin(); }
    const_iterator begin() const { return cells.begin(); }
    iterator       end() { return cells.begin() + std::ptrdiff_t(live()); }
    const_iterator end() const { return cells.begin() + std::ptrdiff_t(live()); }

    void resize(const size_t rows, const size_t cols)
    {
        nRow = rows;
        nCol = cols;
        if (cells.size() < live()) cells = std::vector<T>(live());
    }
    void swap(matrix& other)
    {
        std::swap(nRow, other.nRow);
        std::swap(nCol, other.nCol);
        cells.swap(other.cells);
    }

private:
    size_t         nRow = 0, nCol = 0;
    std::vector<T> cells;

    size_t   live() const { return nRow * nCol; }
    T*       cell(const size_t r, const size_t c) { return cells.data() + (r * nCol + c); }
    const T* cell(const size_t r, const size_t c) const { return cells.data() + (r * nCol + c); }
};

template <class T>
inline matrix<T> transpose(const matrix<T>& in)
{
    matrix<T> out(in.cols(), in.rows());
    for (size_t r = 0; r < in.rows(); ++r) {
        const T* row = in[r];
        for (size_t c = 0; c < in.cols(); ++c) out[c][r] = row[c];
    }
    return out;
}
