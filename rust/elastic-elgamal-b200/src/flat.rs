//! The flat byte layouts of `include/eg_b200.h` as a pair of serde (de)serializers.
//!
//! `Ciphertext`, `RingProof`, `LogEqualityProof` and `VerifiableDecryption` have `to_bytes` forms; `RangeProof`,
//! `QuadraticVotingBallot`, `SumOfSquaresProof`, `CommitmentEquivalenceProof`, `ProofOfPossession` and `EncryptedChoice`
//! only have serde forms.  For a binary (non-human-readable) serializer every element and scalar of those types is one
//! `serialize_bytes` call of 32 bytes (`elastic_elgamal`'s `serde.rs:19-27`), emitted in struct-field order -- which is
//! exactly the concatenation the C ABI documents for each batch entry point.  `to_flat` therefore visits a value and keeps
//! nothing but the byte fields; `from_flat` rebuilds a value from such a concatenation, given the lengths of the sequences
//! it contains in visiting order (a flat layout has no length prefixes: the batch shape fixes them).
use serde::{
    de::{self, DeserializeSeed, SeqAccess, Visitor},
    ser::{self, Impossible},
    Deserialize, Serialize,
};
use std::{collections::VecDeque, fmt};

#[derive(Debug)]
pub struct FlatError(pub String);

impl fmt::Display for FlatError {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        f.write_str(&self.0)
    }
}
impl std::error::Error for FlatError {}
impl ser::Error for FlatError {
    fn custom<T: fmt::Display>(msg: T) -> Self {
        FlatError(msg.to_string())
    }
}
impl de::Error for FlatError {
    fn custom<T: fmt::Display>(msg: T) -> Self {
        FlatError(msg.to_string())
    }
}

/// Appends the flat form of `value` to `out`.
pub fn to_flat<T: Serialize + ?Sized>(value: &T, out: &mut Vec<u8>) -> Result<(), FlatError> {
    value.serialize(FlatSerializer { out })
}

/// Rebuilds a value from its flat form; `seq_lens` are the lengths of the sequences inside it, in visiting order.
pub fn from_flat<'de, T: Deserialize<'de>>(bytes: &'de [u8], seq_lens: &[usize]) -> Result<T, FlatError> {
    let mut de = FlatDeserializer { data: bytes, seq_lens: seq_lens.iter().copied().collect() };
    let value = T::deserialize(&mut de)?;
    if de.data.is_empty() { Ok(value) } else { Err(FlatError(format!("{} trailing bytes", de.data.len()))) }
}

struct FlatSerializer<'a> {
    out: &'a mut Vec<u8>,
}

macro_rules! unsupported {
    ($($name:ident: $ty:ty),*) => {
        $(fn $name(self, _: $ty) -> Result<(), FlatError> { Err(FlatError(concat!(stringify!($name), " has no flat form").into())) })*
    };
}

impl<'a> ser::Serializer for FlatSerializer<'a> {
    type Ok = ();
    type Error = FlatError;
    type SerializeSeq = Self;
    type SerializeTuple = Self;
    type SerializeTupleStruct = Self;
    type SerializeTupleVariant = Impossible<(), FlatError>;
    type SerializeMap = Impossible<(), FlatError>;
    type SerializeStruct = Self;
    type SerializeStructVariant = Impossible<(), FlatError>;

    fn is_human_readable(&self) -> bool {
        false
    }
    fn serialize_bytes(self, v: &[u8]) -> Result<(), FlatError> {
        self.out.extend_from_slice(v);
        Ok(())
    }
    unsupported!(serialize_bool: bool, serialize_i8: i8, serialize_i16: i16, serialize_i32: i32, serialize_i64: i64,
                 serialize_u8: u8, serialize_u16: u16, serialize_u32: u32, serialize_u64: u64, serialize_f32: f32,
                 serialize_f64: f64, serialize_char: char, serialize_str: &str);
    fn serialize_none(self) -> Result<(), FlatError> {
        Ok(())
    }
    fn serialize_some<T: Serialize + ?Sized>(self, value: &T) -> Result<(), FlatError> {
        value.serialize(self)
    }
    fn serialize_unit(self) -> Result<(), FlatError> {
        Ok(())
    }
    fn serialize_unit_struct(self, _: &'static str) -> Result<(), FlatError> {
        Ok(())
    }
    fn serialize_unit_variant(self, _: &'static str, _: u32, _: &'static str) -> Result<(), FlatError> {
        Err(FlatError("enums have no flat form".into()))
    }
    fn serialize_newtype_struct<T: Serialize + ?Sized>(self, _: &'static str, value: &T) -> Result<(), FlatError> {
        value.serialize(self)
    }
    fn serialize_newtype_variant<T: Serialize + ?Sized>(self, _: &'static str, _: u32, _: &'static str, _: &T) -> Result<(), FlatError> {
        Err(FlatError("enums have no flat form".into()))
    }
    fn serialize_seq(self, _: Option<usize>) -> Result<Self, FlatError> {
        Ok(self)
    }
    fn serialize_tuple(self, _: usize) -> Result<Self, FlatError> {
        Ok(self)
    }
    fn serialize_tuple_struct(self, _: &'static str, _: usize) -> Result<Self, FlatError> {
        Ok(self)
    }
    fn serialize_tuple_variant(self, _: &'static str, _: u32, _: &'static str, _: usize) -> Result<Self::SerializeTupleVariant, FlatError> {
        Err(FlatError("enums have no flat form".into()))
    }
    fn serialize_map(self, _: Option<usize>) -> Result<Self::SerializeMap, FlatError> {
        Err(FlatError("maps have no flat form".into()))
    }
    fn serialize_struct(self, _: &'static str, _: usize) -> Result<Self, FlatError> {
        Ok(self)
    }
    fn serialize_struct_variant(self, _: &'static str, _: u32, _: &'static str, _: usize) -> Result<Self::SerializeStructVariant, FlatError> {
        Err(FlatError("enums have no flat form".into()))
    }
}

impl<'a> ser::SerializeSeq for FlatSerializer<'a> {
    type Ok = ();
    type Error = FlatError;
    fn serialize_element<T: Serialize + ?Sized>(&mut self, value: &T) -> Result<(), FlatError> {
        value.serialize(FlatSerializer { out: &mut *self.out })
    }
    fn end(self) -> Result<(), FlatError> {
        Ok(())
    }
}
impl<'a> ser::SerializeTuple for FlatSerializer<'a> {
    type Ok = ();
    type Error = FlatError;
    fn serialize_element<T: Serialize + ?Sized>(&mut self, value: &T) -> Result<(), FlatError> {
        value.serialize(FlatSerializer { out: &mut *self.out })
    }
    fn end(self) -> Result<(), FlatError> {
        Ok(())
    }
}
impl<'a> ser::SerializeTupleStruct for FlatSerializer<'a> {
    type Ok = ();
    type Error = FlatError;
    fn serialize_field<T: Serialize + ?Sized>(&mut self, value: &T) -> Result<(), FlatError> {
        value.serialize(FlatSerializer { out: &mut *self.out })
    }
    fn end(self) -> Result<(), FlatError> {
        Ok(())
    }
}
impl<'a> ser::SerializeStruct for FlatSerializer<'a> {
    type Ok = ();
    type Error = FlatError;
    fn serialize_field<T: Serialize + ?Sized>(&mut self, _: &'static str, value: &T) -> Result<(), FlatError> {
        value.serialize(FlatSerializer { out: &mut *self.out })
    }
    fn end(self) -> Result<(), FlatError> {
        Ok(())
    }
}

struct FlatDeserializer<'de> {
    data: &'de [u8],
    seq_lens: VecDeque<usize>,
}

impl<'de> FlatDeserializer<'de> {
    /// Every byte field of the flat layouts is one 32-byte element or scalar.
    fn take32(&mut self) -> Result<&'de [u8], FlatError> {
        if self.data.len() < 32 {
            return Err(FlatError("flat form is too short".into()));
        }
        let (head, tail) = self.data.split_at(32);
        self.data = tail;
        Ok(head)
    }
}

struct Counted<'a, 'de> {
    de: &'a mut FlatDeserializer<'de>,
    left: usize,
}

impl<'a, 'de> SeqAccess<'de> for Counted<'a, 'de> {
    type Error = FlatError;
    fn next_element_seed<T: DeserializeSeed<'de>>(&mut self, seed: T) -> Result<Option<T::Value>, FlatError> {
        if self.left == 0 {
            return Ok(None);
        }
        self.left -= 1;
        seed.deserialize(&mut *self.de).map(Some)
    }
    fn size_hint(&self) -> Option<usize> {
        Some(self.left)
    }
}

impl<'a, 'de> de::Deserializer<'de> for &'a mut FlatDeserializer<'de> {
    type Error = FlatError;

    fn is_human_readable(&self) -> bool {
        false
    }
    fn deserialize_any<V: Visitor<'de>>(self, _: V) -> Result<V::Value, FlatError> {
        Err(FlatError("the flat form is not self-describing".into()))
    }
    fn deserialize_bytes<V: Visitor<'de>>(self, visitor: V) -> Result<V::Value, FlatError> {
        visitor.visit_borrowed_bytes(self.take32()?)
    }
    fn deserialize_byte_buf<V: Visitor<'de>>(self, visitor: V) -> Result<V::Value, FlatError> {
        visitor.visit_byte_buf(self.take32()?.to_vec())
    }
    fn deserialize_seq<V: Visitor<'de>>(self, visitor: V) -> Result<V::Value, FlatError> {
        let left = self.seq_lens.pop_front().ok_or_else(|| FlatError("missing sequence length".into()))?;
        visitor.visit_seq(Counted { de: self, left })
    }
    fn deserialize_tuple<V: Visitor<'de>>(self, len: usize, visitor: V) -> Result<V::Value, FlatError> {
        visitor.visit_seq(Counted { de: self, left: len })
    }
    fn deserialize_tuple_struct<V: Visitor<'de>>(self, _: &'static str, len: usize, visitor: V) -> Result<V::Value, FlatError> {
        visitor.visit_seq(Counted { de: self, left: len })
    }
    fn deserialize_struct<V: Visitor<'de>>(self, _: &'static str, fields: &'static [&'static str], visitor: V) -> Result<V::Value, FlatError> {
        visitor.visit_seq(Counted { de: self, left: fields.len() })
    }
    fn deserialize_newtype_struct<V: Visitor<'de>>(self, _: &'static str, visitor: V) -> Result<V::Value, FlatError> {
        visitor.visit_newtype_struct(self)
    }
    fn deserialize_unit<V: Visitor<'de>>(self, visitor: V) -> Result<V::Value, FlatError> {
        visitor.visit_unit()
    }
    fn deserialize_unit_struct<V: Visitor<'de>>(self, _: &'static str, visitor: V) -> Result<V::Value, FlatError> {
        visitor.visit_unit()
    }
    fn deserialize_option<V: Visitor<'de>>(self, visitor: V) -> Result<V::Value, FlatError> {
        visitor.visit_some(self)
    }
    serde::forward_to_deserialize_any! {
        bool i8 i16 i32 i64 u8 u16 u32 u64 f32 f64 char str string map enum identifier ignored_any
    }
}
