// cf_matrix.h -- row-major matrix over a std::vector, interface of the reference's matrix<T>
// (matrix.h:9-90): rows(), cols(), operator[] returning the row pointer, resize, iterators,
// conversion between value types, transpose().
#pragma once

#include <algorithm>
#include <cstddef>
#include <vector>

template <class T>
class matrix
{
    size_t         myRows = 0, myCols = 0;
    std::vector<T> myVector;

public:
    matrix() = default;
    matrix(const size_t rows, const size_t cols) : myRows(rows), myCols(cols), myVector(rows * cols) {}
    matrix(const matrix&) = default;
    matrix(matrix&&) noexcept = default;
    matrix& operator=(const matrix&) = default;
    matrix& operator=(matrix&&) noexcept = default;

    template <class U>
    matrix(const matrix<U>& rhs) : myRows(rhs.rows()), myCols(rhs.cols()), myVector(rhs.rows() * rhs.cols())
    {
        auto dst = myVector.begin();
        for (auto src = rhs.begin(); src != rhs.end(); ++src, ++dst) *dst = T(*src);
    }
    template <class U>
    matrix& operator=(const matrix<U>& rhs) { matrix<T> tmp(rhs); swap(tmp); return *this; }

    void swap(matrix& rhs) { myVector.swap(rhs.myVector); std::swap(myRows, rhs.myRows); std::swap(myCols, rhs.myCols); }
    void resize(const size_t rows, const size_t cols)
    {
        myRows = rows; myCols = cols;
        if (myVector.size() < rows * cols) myVector = std::vector<T>(rows * cols);
    }

    size_t rows() const { return myRows; }
    size_t cols() const { return myCols; }
    T* operator[](const size_t row) { return &myVector[row * myCols]; }
    const T* operator[](const size_t row) const { return &myVector[row * myCols]; }
    bool empty() const { return myVector.empty(); }

    using iterator = typename std::vector<T>::iterator;
    using const_iterator = typename std::vector<T>::const_iterator;
    iterator begin() { return myVector.begin(); }
    iterator end() { return myVector.begin() + myRows * myCols; }
    const_iterator begin() const { return myVector.begin(); }
    const_iterator end() const { return myVector.begin() + myRows * myCols; }
};

template <class T>
inline matrix<T> transpose(const matrix<T>& mat)
{
    matrix<T> res(mat.cols(), mat.rows());
    for (size_t i = 0; i < res.rows(); ++i)
        for (size_t j = 0; j < res.cols(); ++j) res[i][j] = mat[j][i];
    return res;
}
